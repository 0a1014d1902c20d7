#!/bin/bash
# round-2 GPU call A: parity after the phase split of the alignment pass, C2 / C3 / NW timings, register-cap A/B
mkdir -p gpurun_out
L=gpurun_out/r02_a.log; : > $L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 >> $L
echo "== pe_perf (C2)" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -2 | cut -c1-400 >> $L
echo "== c3_perf" >> $L
KG_COUNTERS=0 timeout 400 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | cut -c1-1200 >> $L
echo "== nw_perf" >> $L
timeout 300 python tools/nw_perf.py 24000 2>&1 | tail -1 | cut -c1-600 >> $L
for f in "-DAL_MINB_SHORT=8" "-DAL_MINB_SHORT=12" "-DAL_MINB_SHORT=6"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" 2>&1 | grep -A2 "kg_fast_short15aln_pair_kernel" | grep -E "spill|Used" | tr '\n' ' ' >> ../../$L)
  echo >> $L; echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-300 >> $L
done
for f in "-DAL_MINB=4" "-DAL_MINB=8"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" 2>&1 | grep -A2 "kg_fast15aln_pair_kernel" | grep -E "spill|Used" | tr '\n' ' ' >> ../../$L)
  echo >> $L; echo "== variant '$f'" >> $L
  KG_COUNTERS=0 timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_align": [0-9.]*\|"align_gcups": [0-9.]*' | tr '\n' ' ' >> $L
  echo >> $L
done
cat $L
