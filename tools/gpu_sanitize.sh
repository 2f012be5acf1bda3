#!/bin/bash
# compute-sanitizer over the fuzz harness (both decode paths, the unsized decode, the streaming reader): memcheck, racecheck, synccheck
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool" | tee gpurun_out/r02_sanitize_$tool.txt
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/fuzz_gpu.py --count ${COUNT:-150} --streaming 20 2>&1 | grep -v "^=========     \|^=========         " | tail -25 | tee -a gpurun_out/r02_sanitize_$tool.txt
done
