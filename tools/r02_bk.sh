#!/bin/bash
# round-2 GPU call BK: bucket entries {key0, value0, pos, cnt} in place of exist[] -> kv[]: parity, then C2 / C3 / full-scale C5 timings
mkdir -p gpurun_out
L=gpurun_out/r02_bk.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -6 | cut -c1-300 >> $L
echo "== C2 (2 M pairs)" >> $L
timeout 600 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-300 >> $L
echo "== C3 (20 k long reads)" >> $L
KG_COUNTERS=0 timeout 600 python tools/c3_perf.py 20000 0 2>&1 | grep '"mode": "chain"' | cut -c1-330 >> $L
echo "== C5 full scale" >> $L
timeout 900 python tools/c5_perf.py 5000 10000 4000000 2000 2>&1 | tail -3 | cut -c1-420 >> $L
cat $L
