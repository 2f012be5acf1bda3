#!/bin/bash
mkdir -p gpurun_out
BRO_BENCH_MODE=twophase BRO_WORKLOADS=c4_highratio_w16,c5_stored_10k timeout 1200 python tools/quick_perf.py lib_c6.so lib_c2.so 2>&1 | tee -a gpurun_out/quick_variants.log
