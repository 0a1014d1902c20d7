#!/bin/bash
# round-2 GPU call O: warp-aggregated counters, work claimed four tasks at a time, TR_MINB 8: parity + timing, TR_MINB 10 / claim variants
mkdir -p gpurun_out
L=gpurun_out/r02_o.log; : > $L
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_seed.py tests/test_gpu_consensus.py tests/test_gpu_conclave.py -m gpu -q -x --timeout 120 2>&1 | tail -4 >> $L
echo "== pe_perf (C2)" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-400 >> $L
echo "== c2_flow" >> $L
timeout 400 python tools/c2_flow_perf.py 2000000 0 2>&1 | tail -1 | cut -c1-420 >> $L
for f in "-DTR_MINB=10" "-DAL_CLAIM=1 -DKG_CLAIM=1" "-DAL_CLAIM=8 -DKG_CLAIM=8"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu kmagpu_seed.cu && make EXTRA="$f" > /dev/null 2>&1)
  echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
  timeout 300 python tools/c2_flow_perf.py 2000000 0 2>&1 | tail -1 | cut -c90-330 >> $L
done
cat $L
