#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
BRO_BENCH_MODE=twophase BRO_WORKLOADS=c4_highratio_w16,c5b_literals_10k,c3_corpus_x1000 timeout 1200 python tools/quick_perf.py "" lib_allsmem.so lib_litsmem.so 2>&1 | tee -a gpurun_out/quick_variants.log
