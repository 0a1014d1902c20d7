#!/bin/bash
# round-2 GPU call B: parity, per-kernel launch lists of C2 and C3 after the phase split, ncu --set full of the new kernels
mkdir -p gpurun_out
L=gpurun_out/r02_b.log; : > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 >> $L
echo "== pe_perf (C2)" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-400 >> $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r02_b_c2.csv \
    python tools/pe_perf.py 2000000 2 > gpurun_out/ncu_l_c2.log 2>&1
KG_COUNTERS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_b_c3.csv \
    python tools/c3_perf.py 16000 0 > gpurun_out/ncu_l_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"aln_pair_kernel|nw_thread_kernel|aln_emit|aln_reduce|aln_prep" --launch-skip 8 -c 8 \
    -f -o gpurun_out/prof_r02_b_c2 python tools/pe_perf.py 2000000 2 > gpurun_out/ncu_full_b_c2.log 2>&1
KG_COUNTERS=0 ncu --set full --clock-control none --import-source on -k regex:"aln_pair_kernel|nw_thread_kernel|nw_warp_kernel" --launch-skip 12 -c 6 \
    -f -o gpurun_out/prof_r02_b_c3 python tools/c3_perf.py 16000 0 > gpurun_out/ncu_full_b_c3.log 2>&1
tail -2 gpurun_out/ncu_full_b_c3.log >> $L
cat $L
