#!/bin/bash
# fused kernel, second pass (metadata fast path, block-type flags in registers, in-line short copies, long codes inside the
# lane-parallel rounds): GPU suite, fuzz, single-stream times, the fused-kernel workloads
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_quick.log
timeout 600 python tools/fuzz_gpu.py --count 5000 --streaming 60 2>&1 | tail -5 | tee gpurun_out/fuzz_gpu_5000.log
BRO_SINGLE_MODES=fused timeout 300 python tools/time_single.py 20000 2>&1 | tee gpurun_out/time_single.log
BRO_WORKLOADS=c3_corpus_x1000,c1_alice29_single,c2_quickfox_x10k timeout 600 python tools/quick_perf.py "" 2>&1 | tee gpurun_out/quick_product.log
BRO_BENCH_MODE=warp BRO_WORKLOADS=c5b_literals_10k,c6_text_q11_w16,c4_highratio_w16 timeout 600 python tools/quick_perf.py "" 2>&1 | tee -a gpurun_out/quick_product.log
