#!/bin/bash
# round-2 GPU call F: seed kernel occupancy variants (table size x register cap) after the one-round-trip list fetch
mkdir -p gpurun_out
L=gpurun_out/r02_f.log; : > $L
timeout 300 python -m pytest tests/test_gpu_seed.py -m gpu -q --timeout 120 2>&1 | tail -3 >> $L
echo "== base (KG_CAP_LOG=8 KG_LB=1)" >> $L
timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
for f in "-DKG_CAP_LOG=7 -DKG_LB=10" "-DKG_CAP_LOG=7 -DKG_LB=12" "-DKG_CAP_LOG=7 -DKG_LB=8" "-DKG_CAP_LOG=8 -DKG_LB=9"; do
  (cd kma_b200/csrc && touch kmagpu_seed.cu && make EXTRA="$f" 2>&1 | grep -A2 "seed_se_kernelILb0ELb0" | grep -E "spill|Used" | tr '\n' ' ' >> ../../$L)
  echo >> $L; echo "== variant '$f'" >> $L
  KGM=$(echo "$f" | grep -o "KG_LB=[0-9]*" | cut -d= -f2)
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-200 >> $L
done
cat $L
