#!/bin/bash
# round 2, quick check: two-phase parity tests, then kernel-only timing of the given builds on the given workloads
# usage: bash tools/gpu_r2b.sh "<workloads csv>" [lib.so ...]
mkdir -p gpurun_out
WL=${1:-c4_highratio_w16,c5b_literals_10k}; shift
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r2b_pytest.log
BRO_WORKLOADS=$WL timeout 900 python tools/quick_perf.py "" "$@" 2>&1 | tee gpurun_out/r2b_quick.log
