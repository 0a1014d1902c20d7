#!/bin/bash
# round-2 GPU call K (2 GPUs): per-phase wall clock of the whole-span end-to-end step at 1 and 2 ranks
mkdir -p gpurun_out
L=gpurun_out/r02_k.log; : > $L
echo "== 1 rank" >> $L
timeout 300 python tools/e2e_phases.py 2000000 4 2>&1 | grep "^{" | tail -2 >> $L
echo "== 2 ranks" >> $L
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/e2e_phases.py 2000000 4 2>&1 | grep "^{\|rror" | tail -3 >> $L
echo "== host" >> $L
nproc >> $L; grep -c processor /proc/cpuinfo >> $L; cat /sys/fs/cgroup/cpu.max >> $L 2>&1; free -g | head -2 >> $L
cat $L
