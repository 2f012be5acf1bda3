#!/bin/bash
mkdir -p gpurun_out
BRO_BENCH_MODE=twophase BRO_WORKLOADS=c4_highratio_w16,c5b_literals_10k timeout 1200 python tools/quick_perf.py lib_rbc7.so lib_rbc7d5.so 2>&1 | tee -a gpurun_out/quick_variants.log
