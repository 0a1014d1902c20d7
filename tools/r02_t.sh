#!/bin/bash
# round-2 GPU call T: persistent grids sized to one wave, aln_prep / tr_prep as one fused round per word: parity + timing
mkdir -p gpurun_out
L=gpurun_out/r02_t.log; : > $L
timeout 1200 python -m pytest tests/test_gpu_align.py tests/test_gpu_seed.py tests/test_gpu_chain.py tests/test_gpu_consensus.py tests/test_gpu_conclave.py tests/test_gpu_stage1.py -m gpu -q -x --timeout 120 2>&1 | tail -4 >> $L
echo "== pe_perf (C2)" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-400 >> $L
echo "== c2_flow" >> $L
timeout 400 python tools/c2_flow_perf.py 2000000 0 2>&1 | tail -1 | cut -c1-420 >> $L
echo "== c3_perf" >> $L
timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_seed": [0-9.]*\|"ms_align": [0-9.]*\|"ms_reduce": [0-9.]*' | tr '\n' ' ' >> $L
echo >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02_t_flow.csv \
    python tools/c2_flow_perf.py 2000000 0 > gpurun_out/ncu_launch_t.log 2>&1
cat $L
