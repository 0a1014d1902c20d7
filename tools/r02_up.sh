#!/bin/bash
# round-2 GPU call UP: slices of the end-to-end step uploaded in order (one after the other) vs all at once
mkdir -p gpurun_out
L=gpurun_out/r02_up.log; : > $L
for o in 1 0 1; do
  KMA_B200_E2E_UPLOAD_ORDER=$o timeout 200 python bench.py --no-c3 --no-c4 --no-c5 --no-parity --no-cpu-baseline > gpurun_out/bench_up_$o.json 2>/dev/null
  python - <<PY >> $L
import json
d = json.loads([l for l in open("gpurun_out/bench_up_$o.json") if l.startswith("{")][-1])
print("upload order $o: e2e", round(d["e2e"]["value"] / 1e6, 2), "M reads/s,", round(d["e2e"]["ms_per_step"], 1), "ms/step; value", round(d["value"] / 1e6, 1))
PY
done
KMA_B200_E2E_UPLOAD_ORDER=1 timeout 100 python tools/e2e_phases.py 2000000 4 2>/dev/null | tail -1 | cut -c1-900 >> $L
cat $L
