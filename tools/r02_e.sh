#!/bin/bash
# round-2 GPU call E: shim file-level tests, seed kernel after the segment / specialisation changes: timing + ncu --set full
mkdir -p gpurun_out
L=gpurun_out/r02_e.log; : > $L
timeout 900 python -m pytest tests/test_host_shim.py tests/test_gpu_seed.py tests/test_gpu_chain.py tests/test_gpu_conclave.py -m gpu -q --timeout 120 2>&1 | tail -12 >> $L
echo "== pe_perf (C2)" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-400 >> $L
ncu --set full --clock-control none --import-source on -k regex:"seed_se_kernel" --launch-skip 3 -c 1 \
    -f -o gpurun_out/prof_r02_e_seed python tools/pe_perf.py 2000000 2 > gpurun_out/ncu_full_e.log 2>&1
tail -2 gpurun_out/ncu_full_e.log >> $L
for f in "-DKG_MINB=6" "-DKG_MINB=10"; do
  (cd kma_b200/csrc && touch kmagpu_seed.cu && make EXTRA="$f" > /dev/null 2>&1)
  echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
done
cat $L
