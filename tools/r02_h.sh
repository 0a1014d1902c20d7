#!/bin/bash
# round-2 GPU call H: full GPU suite after the column-block NW kernels, NW / C3 / C2 timings
mkdir -p gpurun_out
L=gpurun_out/r02_h.log; : > $L
timeout 1200 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -15 >> $L
echo "== nw_perf" >> $L
timeout 300 python tools/nw_perf.py 24000 2>&1 | tail -1 | cut -c1-700 >> $L
echo "== c3_perf" >> $L
KG_COUNTERS=0 timeout 400 python tools/c3_perf.py 16000 0 2>&1 | tail -2 | grep -o '"mode": "[a-z0-9]*"\|"ms_seed": [0-9.]*\|"ms_align": [0-9.]*\|"align_gcups": [0-9.]*\|"nw_full_cells": [0-9]*\|"nw_band_cells": [0-9]*' | tr '\n' ' ' >> $L
echo >> $L
echo "== pe_perf (C2)" >> $L
timeout 400 python tools/pe_perf.py 2000000 3 2>&1 | tail -1 | cut -c1-300 >> $L
KG_COUNTERS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_h_c3.csv \
    python tools/c3_perf.py 16000 0 > gpurun_out/ncu_l_c3.log 2>&1
cat $L
