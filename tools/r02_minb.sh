#!/bin/bash
# round-2 GPU call MINB: resident-CTA sweep of the short-read pair kernel after the phase split (6 = 80 registers, 8 = 64, 10 = 48)
mkdir -p gpurun_out
L=gpurun_out/r02_minb.log; : > $L
for m in 8 6 10; do
  (cd kma_b200/csrc && touch kmagpu_align.cu kmagpu_align_fast_short.cu && make -j4 EXTRA="-DAL_MINB_SHORT=$m" 2>&1 | grep -A2 "kg_fast_short15aln_pair_kernelILi" | grep "spill\|Used" | tr '\n' ' ' >> ../../$L)
  echo >> $L
  echo "== AL_MINB_SHORT=$m" >> $L
  timeout 300 python tools/pe_perf.py 2000000 4 2>&1 | tail -1 | cut -c1-260 >> $L
done
cat $L
