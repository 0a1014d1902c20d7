#!/bin/bash
# Which part of the NW row sweep costs the alignment kernel: band / full / cells-per-lane variants.
mkdir -p gpurun_out
L=gpurun_out/ab_rs2.log; : > $L
for f in "-DNW_RS_FULL=0" "-DNW_RS_BAND=0" "-DNW_RS_MAXC=2" "-DNW_RS_FULL=0 -DNW_RS_MAXC=4" "-DNW_RS_BAND=0 -DNW_RS_MAXC=1"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" > /dev/null 2>&1)
  echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
  timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"nw_steps": [0-9]*\|"ms_align": [0-9.]*' | tr '\n' ' ' >> $L
  echo >> $L
done
cat $L
