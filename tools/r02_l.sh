#!/bin/bash
# round-2 GPU call L: seed kernel with the segment queue + packed application, merged quick check: parity tests, timing, ncu
mkdir -p gpurun_out
L=gpurun_out/r02_l.log; : > $L
timeout 900 python -m pytest tests/test_gpu_seed.py tests/test_gpu_chain.py tests/test_gpu_align.py tests/test_gpu_conclave.py -m gpu -q -x --timeout 120 2>&1 | tail -12 >> $L
echo "== pe_perf (C2)" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-400 >> $L
echo "== c5_perf" >> $L
timeout 600 python tools/c5_perf.py 2>&1 | tail -2 | cut -c1-600 >> $L
echo "== c3_perf" >> $L
timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_seed": [0-9.]*\|"ms_align": [0-9.]*' | tr '\n' ' ' >> $L
echo >> $L
ncu --set full --clock-control none --import-source on -k regex:"seed_se_kernel" --launch-skip 3 -c 1 \
    -f -o gpurun_out/prof_r02_l_seed python tools/pe_perf.py 2000000 2 > gpurun_out/ncu_full_l.log 2>&1
tail -2 gpurun_out/ncu_full_l.log >> $L
cat $L
