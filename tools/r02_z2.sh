#!/bin/bash
# round-2 GPU call Z2: the -eq / -mi failures of call Z with a record-level message; multi-line FASTA on the device
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stage1.py -m gpu -q --timeout 240 2>&1 | grep -v "^$" | cut -c1-700 | tail -60 > gpurun_out/r02_z2.log
cat gpurun_out/r02_z2.log
