#!/bin/bash
# resident CTAs per SM of the alignment kernels (register cap 65536 / (128 * AL_MINB)): C2 pair kernel and C3 alignment kernel
mkdir -p gpurun_out
L=gpurun_out/ab_minb2.log; : > $L
for f in "-DAL_MINB=6" "-DAL_MINB=8" "-DAL_MINB=10" "-DAL_MINB=12"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" 2>&1 | grep -A2 "aln_pair_kernel" | grep -E "spill|Used" | tr '\n' ' ' >> ../../$L)
  echo >> $L; echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
  timeout 300 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_align": [0-9.]*' | tr '\n' ' ' >> $L
  echo >> $L
done
(cd kma_b200/csrc && touch kmagpu_align.cu && make > /dev/null 2>&1)
cat $L
