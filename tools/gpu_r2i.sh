#!/bin/bash
# fused kernel with lane-parallel literal chunks in the general loop (gpurun -- 'bash tools/gpu_r2i.sh'): GPU suite, fuzz harness,
# kernel-only timings of the fused-kernel workloads
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_quick.log
timeout 300 python tools/fuzz_gpu.py --count 400 --streaming 20 2>&1 | tail -5 | tee gpurun_out/fuzz_gpu.log
BRO_WORKLOADS=c3_corpus_x1000,c1_alice29_single,c2_quickfox_x10k,c4_highratio_w16 timeout 600 python tools/quick_perf.py "" 2>&1 | tee gpurun_out/quick_product.log
BRO_BENCH_MODE=warp BRO_WORKLOADS=c5b_literals_10k,c6_text_q11_w16 timeout 600 python tools/quick_perf.py "" 2>&1 | tee -a gpurun_out/quick_product.log
