#!/bin/bash
# round-2 GPU call S: checkpoint of the bench lines (both arms) + the launch list of the step + ncu of the streaming kernels of the span
mkdir -p gpurun_out
L=gpurun_out/r02_s.log; : > $L
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02_s.json 2> gpurun_out/bench_ref_r02_s.err
timeout 1200 python bench.py > gpurun_out/bench_r02_s.json 2> gpurun_out/bench_r02_s.err
tail -c 300 gpurun_out/bench_r02_s.err >> $L
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_s.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c3 --no-c4 --no-c5 --no-parity > gpurun_out/ncu_launch_s.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"aln_emit_kernel|aln_prep_kernel|tr_prep_kernel|s1_emit_kernel" --launch-skip 4 -c 4 \
    -f -o gpurun_out/prof_r02_s_stream python tools/c2_flow_perf.py 2000000 0 > gpurun_out/ncu_full_s.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_r02_s_stream.ncu-rep gpurun_out/ncu_r02_s_stream.txt >> $L 2>&1
ls -la gpurun_out/*.ncu-rep >> $L
python - <<'PY' >> $L 2>&1
import json
d = json.loads([l for l in open("gpurun_out/bench_r02_s.json") if l.startswith("{")][-1])
print("value", round(d["value"] / 1e6, 2), "e2e", round(d["e2e"]["value"] / 1e6, 2), "ms", round(d["e2e"]["ms_per_step"], 1), "hot", round(d["e2e_hotpath"]["value"] / 1e6, 2), d["stage_ms"])
print("nw", d["nw"]["gcups"], "c3", d["c3"]["reads_per_s"], d["c3"]["align_gcups"], "c4", d["c4"]["reads_per_s"], "c5", d["c5"]["reads_per_s"], d["c5"]["seed_kernel_ms"])
print("parity", {k: d["parity"].get(k) for k in ("scores_equal", "frag_sorted_equal")}, d["parity"].get("files"), "cpu", d["cpu_baseline"]["value"])
r = json.loads([l for l in open("gpurun_out/bench_ref_r02_s.json") if l.startswith("{")][-1])
print("reference arm", r["value"], r["ms_per_step"])
PY
cat $L
