#!/bin/bash
# One GPU-box call with the round's evidence: every GPU parity test, smoke, both bench arms, the ncu launch list of the bench step,
# ncu --set full of the dominant kernels of C2 / C3 / C5 / stand-alone NW. usage: bash tools/r02_final.sh <tag>
tag=${1:-r02_v5}
mkdir -p gpurun_out
L=gpurun_out/final_$tag.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -3 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log >> $L
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
timeout 1200 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 400 gpurun_out/bench_$tag.err >> $L
cut -c1-300 gpurun_out/bench_$tag.json >> $L
timeout 900 python bench.py --e2e-workers 8 --no-cpu-baseline --no-c3 --no-c4 --no-c5 --no-parity > gpurun_out/bench_${tag}_w8.json 2>/dev/null
python - <<PY >> $L 2>&1
import json
for f in ("gpurun_out/bench_$tag.json", "gpurun_out/bench_${tag}_w8.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f, "value", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2), "ms", round(d["e2e"]["ms_per_step"], 1), "hot", round(d["e2e_hotpath"]["value"] / 1e6, 2))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c3 --no-c4 --no-c5 --no-parity > gpurun_out/ncu_launch_$tag.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"aln_pair_kernel|seed_se_kernel|nw_thread_kernel" --launch-skip 12 -c 7 \
    -f -o gpurun_out/prof_${tag}_c2 python tools/pe_perf.py 2000000 5 > gpurun_out/ncu_full_${tag}_c2.log 2>&1
tail -1 gpurun_out/ncu_full_${tag}_c2.log >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"aln_pair_kernel|chain_kernel|nw_warp_kernel|nw_thread_kernel" --launch-skip 12 -c 6 \
    -f -o gpurun_out/prof_${tag}_c3 env KG_COUNTERS=0 python tools/c3_perf.py 20000 0 > gpurun_out/ncu_full_${tag}_c3.log 2>&1
tail -1 gpurun_out/ncu_full_${tag}_c3.log | cut -c1-200 >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"seed_se_kernel" --launch-skip 3 -c 1 \
    -f -o gpurun_out/prof_${tag}_c5 python tools/c5_perf.py 5000 10000 4000000 0 > gpurun_out/ncu_full_${tag}_c5.log 2>&1
tail -1 gpurun_out/ncu_full_${tag}_c5.log | cut -c1-200 >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nw_warp_kernel" --launch-skip 2 -c 1 \
    -f -o gpurun_out/prof_${tag}_nw python tools/nw_perf.py 24000 > gpurun_out/ncu_full_${tag}_nw.log 2>&1
tail -1 gpurun_out/ncu_full_${tag}_nw.log | cut -c1-200 >> $L
# the reports stay on the box (together they exceed what comes back): text summaries, per-launch DRAM traffic and the
# per-instruction tables of the two big C2 kernels come back instead
for x in c2 c3 c5 nw; do
  units=2000000; [ $x = c3 ] && units=20000; [ $x = c5 ] && units=4000000; [ $x = nw ] && units=24000
  python tools/ncu_summary.py gpurun_out/prof_${tag}_$x.ncu-rep gpurun_out/ncu_${tag}_$x.txt gpurun_out/traffic_${tag}_$x.json $units >> $L 2>&1
done
ncu -i gpurun_out/prof_${tag}_c2.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/sass_${tag}_c2.csv.gz
ncu -i gpurun_out/prof_${tag}_c3.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip -9 > gpurun_out/sass_${tag}_c3.csv.gz
python tools/merge_traffic.py $tag >> $L 2>&1
rm -f gpurun_out/prof_${tag}_*.ncu-rep
du -sh gpurun_out >> $L
cat $L
