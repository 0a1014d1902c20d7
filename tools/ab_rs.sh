#!/bin/bash
# A/B of the NW row sweep against the wavefront on the GPU box: parity first, then PE / C3 / NW-only timings.
mkdir -p gpurun_out
L=gpurun_out/ab_rs.log; : > $L
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_chain.py -x -q 2>&1 | tail -5 >> $L
for f in "" "-DKG_NO_RS"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" > /dev/null 2>&1)
  echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-300 >> $L
  timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | cut -c1-900 >> $L
  timeout 300 python tools/nw_perf.py 6000 2>&1 | tail -1 | cut -c1-600 >> $L
done
cat $L
