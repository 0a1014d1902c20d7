#!/bin/bash
# round-2 GPU call M: the whole span (FASTQ text -> consensus) on one handle: stage wall clock, ncu launch list, ncu --set full of the
# traceback kernel
mkdir -p gpurun_out
L=gpurun_out/r02_m.log; : > $L
timeout 400 python tools/c2_flow_perf.py 2000000 20000 2>&1 | tail -2 | cut -c1-700 >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_m_flow.csv \
    python tools/c2_flow_perf.py 2000000 0 > gpurun_out/ncu_launch_m.log 2>&1
tail -1 gpurun_out/ncu_launch_m.log | cut -c1-300 >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tr_task_kernel|tr_matrix_kernel" --launch-skip 2 -c 2 \
    -f -o gpurun_out/prof_r02_m_trace python tools/c2_flow_perf.py 2000000 0 > gpurun_out/ncu_full_m.log 2>&1
tail -2 gpurun_out/ncu_full_m.log >> $L
cat $L
