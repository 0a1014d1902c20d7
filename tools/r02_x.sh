#!/bin/bash
# round-2 GPU call X: GPU parity incl. -lc in the chain scan, the frag_raw writer without local memory, C5 at full scale (50k templates / 500 Mb)
mkdir -p gpurun_out
L=gpurun_out/r02_x.log; : > $L
timeout 1200 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -5 >> $L
( time timeout 1500 python bench.py --no-cpu-baseline --no-c3 --no-c4 --no-parity > gpurun_out/bench_r02_x.json 2> gpurun_out/bench_r02_x.err ) 2>> $L
tail -c 600 gpurun_out/bench_r02_x.err >> $L
python - <<PY >> $L 2>&1
import json
d = json.loads([l for l in open("gpurun_out/bench_r02_x.json") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "stage", d["stage_ms"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("c5", json.dumps(d.get("c5"))[:1500])
print("side_errors", d.get("side_errors"))
PY
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> $L
free -g | head -2 >> $L
cat $L
