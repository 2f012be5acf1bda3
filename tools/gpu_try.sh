#!/bin/bash
# Development round on a GPU box: parity tests, kernel-only timings per build variant / knob.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
BRO_BENCH_MODE=twophase timeout 900 python tools/quick_perf.py "" 2>&1 | tee -a gpurun_out/quick_twophase.log
BRO_WORKLOADS=c2_quickfox_x10k,c3_corpus_x1000 timeout 900 python tools/quick_perf.py "" 2>&1 | tee -a gpurun_out/quick_auto.log
