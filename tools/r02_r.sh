#!/bin/bash
# round-2 GPU call R (2 GPUs): bench line at N = 2 after the consensus-download and kernel changes; phase timeline at 2 ranks
mkdir -p gpurun_out
L=gpurun_out/r02_r.log; : > $L
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_r_2gpu.json 2> gpurun_out/bench_r02_r_2gpu.err ) 2>> $L
tail -c 600 gpurun_out/bench_r02_r_2gpu.err >> $L
python - <<'PY' >> $L 2>&1
import json
d = json.loads([l for l in open("gpurun_out/bench_r02_r_2gpu.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "n_gpus", "ms_per_step")})
print("e2e", {k: d["e2e"][k] for k in ("value", "ms_per_step", "host_link_GBs")})
print("hot", d["e2e_hotpath"]["value"], "allreduce ms", d["config"]["allreduce_scores_ms"], d["config"]["allreduce_matrix_ms"])
print("c5", {k: d.get("c5", {}).get(k) for k in ("reads_per_s", "ms", "seed_kernel_ms")}, d.get("side_errors"))
PY
echo "== 2 ranks phases" >> $L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/e2e_phases.py 2000000 4 2>&1 | grep "^{" | tail -2 | cut -c1-1500 >> $L
cat $L
