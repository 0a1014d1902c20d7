#!/bin/bash
# NwScratch passed to nw_warp by value (this tree) vs by reference (sed back): production pair kernel on C2 and C3
mkdir -p gpurun_out
L=gpurun_out/ab_byval.log; : > $L
for v in byvalue byref; do
  if [ $v = byref ]; then sed -i 's/int band, const NwScratch ws, NwStat \*out,/int band, const NwScratch \&ws, NwStat *out,/' kma_b200/csrc/kmagpu_nw.cuh; fi
  (cd kma_b200/csrc && touch kmagpu_align.cu kmagpu_align_fast.cu && make -j4 > /dev/null 2>&1)
  echo "== $v" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
  KG_COUNTERS=0 timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_align": [0-9.]*' | tr '\n' ' ' >> $L
  echo >> $L
done
cat $L
