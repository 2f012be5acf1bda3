#!/bin/bash
# Development round on a GPU box: parity tests, kernel-only timings per build variant, ncu captures of the two-phase kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
BRO_BENCH_MODE=twophase timeout 900 python tools/quick_perf.py 2>&1 | tee gpurun_out/quick_twophase.log
BRO_BENCH_MODE=twophase BRO_WORKLOADS=c4_highratio_w16 timeout 900 python tools/quick_perf.py lib_p6.so lib_d2.so 2>&1 | tee -a gpurun_out/quick_twophase.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bro_parse_kernel|bro_copy_kernel" -s 2 -c 2 -f -o gpurun_out/prof_c4_twophase \
    python bench.py --mode twophase --streams 20000 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_c4_twophase.log 2>&1
tail -2 gpurun_out/prof_c4_twophase.log | cut -c1-200
