#!/bin/bash
# round 2, call A: baseline check of the round-1 build + the staged copy-kernel variants (kernel-only timing)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee gpurun_out/r2a_gpu.txt
BRO_WORKLOADS=c4_highratio_w16,c5_stored_10k timeout 900 python tools/quick_perf.py "" lib_st3.so lib_st4.so lib_st5.so 2>&1 | tee gpurun_out/r2a_quick_variants.log
for v in lib_st4.so; do
  BRO_B200_LIB=$PWD/brotli_rs_b200/lib/$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "twophase or auto or side_by_side" 2>&1 | tail -3 | tee -a gpurun_out/r2a_pytest_variant.log
done
