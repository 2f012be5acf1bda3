#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/fuzz_gpu.py --count 3000 --seed 5 > gpurun_out/fuzz_gpu.log 2>&1; tail -4 gpurun_out/fuzz_gpu.log
timeout 700 compute-sanitizer --tool memcheck --print-limit 20 python tools/fuzz_gpu.py --count 600 --seed 9 > gpurun_out/fuzz_memcheck.log 2>&1
grep "ERROR SUMMARY\|fuzz_gpu:\|Invalid" gpurun_out/fuzz_memcheck.log | head -8
