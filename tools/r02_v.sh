#!/bin/bash
# round-2 GPU call V: resident-CTA sweep of the short-read pair kernel after the phase split; ncu of the pair and traceback kernels
mkdir -p gpurun_out
L=gpurun_out/r02_v.log; : > $L
for f in "-DAL_MINB_SHORT=8" "-DAL_MINB_SHORT=12" "-DAL_MINB_SHORT=10"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu kmagpu_align_fast_short.cu kmagpu_align_fast.cu && make EXTRA="$f" 2>&1 | grep -A2 "kg_fast_short15aln_pair_kernelILi" | grep "spill\|Used" | tr '\n' ' ' >> ../../$L)
  echo >> $L
  echo "== variant '$f'" >> $L
  timeout 300 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-220 >> $L
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"aln_pair_kernel|tr_task_kernel" --launch-skip 2 -c 2 \
    -f -o gpurun_out/prof_r02_v python tools/c2_flow_perf.py 2000000 0 > gpurun_out/ncu_full_v.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_v.ncu-rep gpurun_out/ncu_r02_v.txt >> $L 2>&1
ls -la gpurun_out/prof_r02_v.ncu-rep >> $L
cat $L
