#!/bin/bash
# resident CTAs per SM of the seeding kernel
mkdir -p gpurun_out
L=gpurun_out/ab_seed.log; : > $L
for f in "-DKG_MINB=8" "-DKG_MINB=9" "-DKG_MINB=6"; do
  (cd kma_b200/csrc && touch kmagpu_seed.cu && make EXTRA="$f" 2>&1 | grep -A3 "seed_se_kernelILb0" | grep -E "spill|Used" | tr '\n' ' ' >> ../../$L)
  echo >> $L; echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
done
(cd kma_b200/csrc && touch kmagpu_seed.cu && make > /dev/null 2>&1)
cat $L
