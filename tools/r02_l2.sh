#!/bin/bash
# round-2 GPU call L2: cudaLimitMaxL2FetchGranularity 32 / 64 / 128 on the random gathers -- full-scale C5 (seed + alignment pass), C2, C3
mkdir -p gpurun_out
L=gpurun_out/r02_l2.log; : > $L
for g in 64 32 128 32; do
  echo "== C5 full scale, L2 fetch $g" >> $L
  KMAGPU_L2_FETCH=$g timeout 900 python tools/c5_perf.py 5000 10000 4000000 0 2>&1 | tail -1 | cut -c1-420 >> $L
done
for g in 64 32; do
  echo "== C2 (2 M pairs), L2 fetch $g" >> $L
  KMAGPU_L2_FETCH=$g timeout 600 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-300 >> $L
  echo "== C3 (20 k long reads), L2 fetch $g" >> $L
  KMAGPU_L2_FETCH=$g KG_COUNTERS=0 timeout 600 python tools/c3_perf.py 20000 0 2>&1 | grep '"mode": "chain"' | cut -c1-330 >> $L
done
cat $L
