#!/bin/bash
# Development round on a GPU box (gpurun -- 'bash tools/gpu_try.sh'): parity tests, then kernel-only timings of the
# headline workloads for the product build (tools/quick_perf.py takes variant libraries built by tools/build_variants.py).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
BRO_WORKLOADS=c4_highratio_w16,c5b_literals_10k,c5_stored_10k timeout 1200 python tools/quick_perf.py "" 2>&1 | tee gpurun_out/quick_auto.log
