#!/bin/bash
# round-2 GPU call U: small thread classes through the warp kernel; resident-CTA sweep of the record-streaming kernels
mkdir -p gpurun_out
L=gpurun_out/r02_u.log; : > $L
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_consensus.py tests/test_gpu_conclave.py -m gpu -q -x --timeout 120 2>&1 | tail -3 >> $L
echo "== pe_perf (ST_MINB 1)" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-220 >> $L
timeout 400 python tools/c2_flow_perf.py 2000000 0 2>&1 | tail -1 | cut -c90-330 >> $L
for f in "-DST_MINB=8" "-DST_MINB=6"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" > /dev/null 2>&1)
  echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-220 >> $L
  timeout 300 python tools/c2_flow_perf.py 2000000 0 2>&1 | tail -1 | cut -c90-330 >> $L
done
cat $L
