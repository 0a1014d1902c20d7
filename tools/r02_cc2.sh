#!/bin/bash
# round-2 GPU call CC2: ConClave 2 on the device + the whole GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 240 2>&1 | grep -v "^$" | cut -c1-600 | tail -40 > gpurun_out/r02_cc2.log
cat gpurun_out/r02_cc2.log
