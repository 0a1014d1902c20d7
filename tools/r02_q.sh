#!/bin/bash
# round-2 GPU call Q: row sweep with the lean interior rows: parity, stand-alone GCUPS, C3; resident-CTA variants
mkdir -p gpurun_out
L=gpurun_out/r02_q.log; : > $L
timeout 900 python -m pytest tests/test_gpu_align.py -m gpu -q -x --timeout 120 2>&1 | tail -3 >> $L
echo "== nw_perf (NW_NARROW_MINB 6)" >> $L
timeout 300 python tools/nw_perf.py 24000 2>&1 | tail -1 | cut -c150-330 >> $L
echo "== c3_perf" >> $L
timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_seed": [0-9.]*\|"ms_align": [0-9.]*\|"align_gcups": [0-9.]*' | tr '\n' ' ' >> $L
echo >> $L
for f in "-DNW_NARROW_MINB=5" "-DNW_NARROW_MINB=4"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" > /dev/null 2>&1)
  echo "== variant '$f'" >> $L
  timeout 300 python tools/nw_perf.py 24000 2>&1 | tail -1 | cut -c150-330 >> $L
done
cat $L
