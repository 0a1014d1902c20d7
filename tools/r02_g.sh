#!/bin/bash
# round-2 GPU call G: the new bench line (parity legs, C5, whole-span e2e) + reference arm
mkdir -p gpurun_out
L=gpurun_out/r02_g.log; : > $L
( time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_g.json 2> gpurun_out/bench_r02_g.err ) 2>> $L
tail -c 1500 gpurun_out/bench_r02_g.err >> $L
python - <<'PY' >> $L 2>&1
import json
d = json.load(open("gpurun_out/bench_r02_g.json"))
def show(k, v, ind=0):
    if isinstance(v, dict):
        print(" " * ind + k + ":")
        for kk, vv in v.items():
            show(kk, vv, ind + 2)
    else:
        t = str(v)
        print(" " * ind + f"{k}: {t[:140]}")
for k in ("value", "ms_per_step", "e2e", "e2e_hotpath", "stage_ms", "parity", "roofline", "nw", "side_errors", "cpu_baseline"):
    if k in d: show(k, d[k])
for k in ("c3", "c4", "c5"):
    if k in d:
        show(k, {kk: vv for kk, vv in d[k].items() if kk in ("reads_per_s", "ms", "parity", "roofline", "cpu_reference", "align_gcups", "seed_kernel_ms", "align_ms", "db", "error")})
PY
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r02_g.json 2> gpurun_out/bench_ref_r02_g.err ) 2>> $L
cut -c1-300 gpurun_out/bench_ref_r02_g.json >> $L
cat $L
