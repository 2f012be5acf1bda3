#!/bin/bash
# Development round on a GPU box (gpurun -- 'bash tools/gpu_try.sh lib_a.so lib_b.so ...'): kernel-only timings of the
# headline workloads for variant libraries built by tools/build_variants.py (bench.py checks statuses, lengths and
# sampled slots of every batch it times), then the parity tests selected by BRO_TRY_TESTS on the product build.
mkdir -p gpurun_out
if [ -n "$BRO_TRY_TESTS" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$BRO_TRY_TESTS" 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_try.log
fi
if [ $# -gt 0 ]; then
  BRO_WORKLOADS=${BRO_WORKLOADS:-c4_highratio_w16} timeout 600 python tools/quick_perf.py "$@" 2>&1 | tee gpurun_out/quick_variants.log
fi
