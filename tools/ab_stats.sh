#!/bin/bash
# statistic counters of the alignment kernels compiled in / out
mkdir -p gpurun_out
L=gpurun_out/ab_stats.log; : > $L
for f in "" "-DKG_NO_STATS"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" > /dev/null 2>&1)
  echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
  timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_align": [0-9.]*' | tr '\n' ' ' >> $L
  echo >> $L
done
(cd kma_b200/csrc && touch kmagpu_align.cu && make > /dev/null 2>&1)
cat $L
