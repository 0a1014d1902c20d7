#!/bin/bash
# round-2 GPU call Z: GPU parity incl. -eq / -mi of stage 1; ncu --set full of the main seed kernel build on C2 and on the full-scale C5
# (call r02_v3 caught the two small passes instead: three seed_se_kernel launches per step since the generic / dense passes run every time)
tag=r02_v3
mkdir -p gpurun_out
L=gpurun_out/r02_z.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -5 >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"seed_se_kernel" --launch-skip 3 -c 3 \
    -f -o gpurun_out/prof_${tag}_c2seed python tools/pe_perf.py 2000000 2 > gpurun_out/ncu_full_${tag}_c2seed.log 2>&1
tail -1 gpurun_out/ncu_full_${tag}_c2seed.log | cut -c1-200 >> $L
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"seed_se_kernel" --launch-skip 3 -c 3 \
    -f -o gpurun_out/prof_${tag}_c5seed python tools/c5_perf.py 5000 10000 4000000 0 > gpurun_out/ncu_full_${tag}_c5seed.log 2>&1
tail -1 gpurun_out/ncu_full_${tag}_c5seed.log | cut -c1-200 >> $L
python tools/ncu_summary.py gpurun_out/prof_${tag}_c2seed.ncu-rep gpurun_out/ncu_${tag}_c2seed.txt gpurun_out/traffic_${tag}_c2seed.json 2000000 >> $L 2>&1
python tools/ncu_summary.py gpurun_out/prof_${tag}_c5seed.ncu-rep gpurun_out/ncu_${tag}_c5seed.txt gpurun_out/traffic_${tag}_c5seed.json 4000000 >> $L 2>&1
ncu -i gpurun_out/prof_${tag}_c5seed.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/sass_${tag}_c5seed.csv.gz
rm -f gpurun_out/prof_${tag}_*.ncu-rep
cat $L
