#!/bin/bash
mkdir -p gpurun_out
BRO_WORKLOADS=c4_highratio_w16,c5b_literals_10k timeout 1200 python tools/quick_perf.py "" 2>&1 | tee -a gpurun_out/quick_auto.log
