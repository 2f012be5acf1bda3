#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
BRO_WORKLOADS=c4_highratio_w16,c5_stored_10k timeout 1200 python tools/quick_perf.py "" 2>&1 | tee -a gpurun_out/quick_variants.log
