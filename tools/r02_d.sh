#!/bin/bash
# round-2 GPU call D: the hanging shim case with traces, then the rest of the shim tests with a per-test timeout
mkdir -p gpurun_out
L=gpurun_out/r02_d.log; : > $L
mkdir -p /tmp/sd && cd /tmp/sd && python - <<'PY' >> $GRAFT_REPO_ROOT/$L 2>&1
import sys, pathlib
sys.path.insert(0, "/root/repo")
from tests import test_host_shim as t
tmp = pathlib.Path("/tmp/sd")
args = t._make_case(tmp, "c2_pe_apm_p")
open("/tmp/sd/args.txt", "w").write(" ".join(args))
PY
cd /tmp/sd && KMAGPU_DEBUG=1 timeout 60 /root/repo/oracle/_ref/kma_gpu $(cat args.txt) -o gpu -t 1 > /tmp/sd/out.txt 2>&1; echo "rc=$?" >> $GRAFT_REPO_ROOT/$L; tail -30 /tmp/sd/out.txt >> $GRAFT_REPO_ROOT/$L
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_host_shim.py tests/test_gpu_seed.py tests/test_gpu_align.py -m gpu -q --timeout 90 2>&1 | tail -25 >> $L
echo "== pe_perf (C2) with the segment seed kernel" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-400 >> $L
cat $L
