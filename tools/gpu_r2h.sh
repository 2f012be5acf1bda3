#!/bin/bash
# kernel-only timings of the product build on every workload + the two-phase parity tests (gpurun -- 'bash tools/gpu_r2h.sh [libs]')
mkdir -p gpurun_out
BRO_WORKLOADS=${BRO_WORKLOADS:-c4_highratio_w16,c5_stored_10k,c5b_literals_10k,c7_far_w22,c6_text_q11_w16} timeout 600 python tools/quick_perf.py "" "$@" 2>&1 | tee gpurun_out/quick_product.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_quick.log
