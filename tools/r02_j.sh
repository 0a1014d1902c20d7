#!/bin/bash
# round-2 GPU call J (2 GPUs): the exchanges through the C entry points (tools/dist_check.py) and the bench line at N = 2
mkdir -p gpurun_out
L=gpurun_out/r02_j.log; : > $L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -12 >> $L
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_j_2gpu.json 2> gpurun_out/bench_r02_j_2gpu.err ) 2>> $L
tail -c 1200 gpurun_out/bench_r02_j_2gpu.err >> $L
python - <<'PY' >> $L 2>&1
import json
d = json.loads([l for l in open("gpurun_out/bench_r02_j_2gpu.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")})
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "host_link_GBs")})
print("allreduce ms", d["config"]["allreduce_scores_ms"], d["config"]["allreduce_matrix_ms"])
print("c5", {k: d.get("c5", {}).get(k) for k in ("reads_per_s", "ms", "seed_kernel_ms")}, d.get("side_errors"))
PY
cat $L
