#!/bin/bash
# One GPU-box call: parity tests, both bench arms, ncu launch list, ncu --set full of the two dominant kernels.
# usage (from the repo root on the box): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_$tag.log
cat gpurun_out/pytest_$tag.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -c 600 gpurun_out/bench_$tag.err
cut -c1-400 gpurun_out/bench_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-c3 --no-c4 --no-text > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"aln_pair_kernel|seed_se_kernel" --launch-skip 2 -c 2 \
    -f -o gpurun_out/prof_$tag python tools/pe_perf.py 2000000 2 > gpurun_out/ncu_full_$tag.log 2>&1
tail -3 gpurun_out/ncu_full_$tag.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1; tail -1 gpurun_out/smoke_$tag.log
ncu --set full --clock-control none --import-source on -k regex:"aln_pair_kernel|chain_kernel" --launch-skip 4 -c 2 \
    -f -o gpurun_out/prof_c3_$tag python tools/c3_perf.py 16000 0 > gpurun_out/ncu_c3_$tag.log 2>&1
tail -2 gpurun_out/ncu_c3_$tag.log | cut -c1-200
