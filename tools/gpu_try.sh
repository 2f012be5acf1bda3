#!/bin/bash
mkdir -p gpurun_out
BRO_BENCH_MODE=twophase BRO_WORKLOADS=c4_highratio_w16,c5b_literals_10k timeout 1200 python tools/quick_perf.py lib_u16.so lib_u32.so lib_pat256.so 2>&1 | tee -a gpurun_out/quick_variants.log
