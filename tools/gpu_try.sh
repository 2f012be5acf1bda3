#!/bin/bash
# Development round on a GPU box: parity tests, then kernel-only timings of every BASELINE workload per path.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
BRO_BENCH_MODE=twophase timeout 900 python tools/quick_perf.py 2>&1 | tee gpurun_out/quick_twophase.log
BRO_BENCH_MODE=warp BRO_WORKLOADS=c4_highratio_w16,c5b_literals_10k timeout 600 python tools/quick_perf.py 2>&1 | tee gpurun_out/quick_warp.log
