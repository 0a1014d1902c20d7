#!/bin/bash
# round-2 GPU call N: traceback kernel (1x1 closed form, shared parameter blocks, anker_rc out of line): parity, timing, resident-CTA sweep
mkdir -p gpurun_out
L=gpurun_out/r02_n.log; : > $L
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_consensus.py tests/test_gpu_conclave.py tests/test_host_shim.py -m gpu -q -x --timeout 120 2>&1 | tail -6 >> $L
echo "== c2_flow (TR_MINB default)" >> $L
timeout 400 python tools/c2_flow_perf.py 2000000 0 2>&1 | tail -1 | cut -c1-420 >> $L
echo "== c4" >> $L
timeout 400 python tools/c4_perf.py 5000000 1000000 0 2>&1 | tail -1 | cut -c1-600 >> $L
for f in "-DTR_MINB=4" "-DTR_MINB=5" "-DTR_MINB=8"; do
  (cd kma_b200/csrc && touch kmagpu_align.cu && make EXTRA="$f" 2>&1 | grep -A2 "tr_task_kernel" | grep "spill\|Used" | tr '\n' ' ' >> ../../$L)
  echo >> $L
  echo "== variant '$f'" >> $L
  timeout 300 python tools/c2_flow_perf.py 2000000 0 2>&1 | tail -1 | cut -c1-420 >> $L
done
cat $L
