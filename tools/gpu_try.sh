#!/bin/bash
# Development round on a GPU box (gpurun -- 'bash tools/gpu_try.sh lib_a.so lib_b.so ...'): kernel-only timings of the
# headline workloads for variant libraries built by tools/build_variants.py (bench.py checks statuses, lengths and
# sampled slots of every batch it times), then the two-phase parity tests on the first variant.
mkdir -p gpurun_out
BRO_WORKLOADS=${BRO_WORKLOADS:-c4_highratio_w16} timeout 600 python tools/quick_perf.py "$@" 2>&1 | tee gpurun_out/quick_variants.log
BRO_WORKLOADS=c5_stored_10k,c5b_literals_10k timeout 200 python tools/quick_perf.py "$1" 2>&1 | tee -a gpurun_out/quick_variants.log
BRO_B200_LIB=$PWD/brotli_rs_b200/lib/$1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "twophase or auto or side_by_side or size_hints or reservation" 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_variant.log
