#!/bin/bash
# round-2 GPU call W: state of HEAD after the container was re-created -- every GPU parity test, smoke, the default bench line
mkdir -p gpurun_out
L=gpurun_out/r02_w.log; : > $L
timeout 1200 python -m pytest tests -m gpu -q --timeout 180 2>&1 | tail -3 >> $L
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 >> $L
timeout 1200 python bench.py > gpurun_out/bench_r02_w.json 2> gpurun_out/bench_r02_w.err
tail -c 600 gpurun_out/bench_r02_w.err >> $L
cut -c1-400 gpurun_out/bench_r02_w.json >> $L
cat $L
