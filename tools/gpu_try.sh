#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 1200 python tools/quick_perf.py "" 2>&1 | tee -a gpurun_out/quick_auto.log
