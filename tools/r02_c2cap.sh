#!/bin/bash
# round-2 GPU call C2CAP: ncu --set full of the C2 step's kernels after the bucket layout (the capture window of r02_final.sh v4 was past the run's end)
tag=r02_v4
mkdir -p gpurun_out
L=gpurun_out/r02_c2cap.log; : > $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"aln_pair_kernel|seed_se_kernel|nw_thread_kernel" --launch-skip 12 -c 7 \
    -f -o gpurun_out/prof_${tag}_c2 python tools/pe_perf.py 2000000 5 > gpurun_out/ncu_full_${tag}_c2.log 2>&1
tail -1 gpurun_out/ncu_full_${tag}_c2.log | cut -c1-200 >> $L
python tools/ncu_summary.py gpurun_out/prof_${tag}_c2.ncu-rep gpurun_out/ncu_${tag}_c2.txt gpurun_out/traffic_${tag}_c2.json 2000000 >> $L 2>&1
ncu -i gpurun_out/prof_${tag}_c2.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/sass_${tag}_c2.csv.gz
rm -f gpurun_out/prof_${tag}_*.ncu-rep
grep "^== " -A1 gpurun_out/ncu_${tag}_c2.txt | cut -c1-120 >> $L
cat $L
