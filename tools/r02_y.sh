#!/bin/bash
# round-2 GPU call Y: GPU parity of -proxi / presets / -ts on long reads; chain kernel time after the -lc / -proxi additions
mkdir -p gpurun_out
L=gpurun_out/r02_y.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q --timeout 240 2>&1 | tail -40 >> $L
timeout 600 env KG_COUNTERS=0 python tools/c3_perf.py 20000 0 2>&1 | tail -3 | cut -c1-600 >> $L
cat $L
