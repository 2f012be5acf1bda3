#!/bin/bash
# Development round on a GPU box (gpurun -- 'bash tools/gpu_try.sh lib_a.so lib_b.so ...'): the parity tests selected by
# BRO_TRY_TESTS on the product build, kernel-only timings of the headline workloads for variant libraries built by
# tools/build_variants.py (bench.py checks statuses, lengths and sampled slots of every batch it times), and with
# BRO_TRY_VARIANT_TESTS=1 the two-phase parity tests on every variant.
mkdir -p gpurun_out
if [ -n "$BRO_TRY_TESTS" ]; then
  timeout 300 python -m pytest tests -m gpu -q -k "$BRO_TRY_TESTS" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_try.log
fi
if [ $# -gt 0 ]; then
  BRO_WORKLOADS=${BRO_WORKLOADS:-c4_highratio_w16} timeout 300 python tools/quick_perf.py "$@" 2>&1 | tee gpurun_out/quick_variants.log
fi
if [ -n "$BRO_TRY_VARIANT_TESTS" ]; then
  for v in "$@"; do
    echo "== $v" | tee -a gpurun_out/pytest_gpu_variant.log
    BRO_B200_LIB=$PWD/brotli_rs_b200/lib/$v timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "twophase" 2>&1 | tail -3 | tee -a gpurun_out/pytest_gpu_variant.log
  done
fi
