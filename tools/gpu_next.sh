#!/bin/bash
# First GPU call of the next round (build the variants here first:
#   python tools/build_variants.py st3=BRO_COPY_STAGED=1 st4=BRO_COPY_STAGED=1,BRO_COPY_MIN_BLOCKS=4 st5=BRO_COPY_STAGED=1,BRO_COPY_MIN_BLOCKS=5
# then  gpurun --timeout 900 -- 'bash tools/gpu_next.sh'):
# the whole GPU suite on the product build (two tests were added after the last box time of round 1: the largest window
# and the corpus driver), the staged copy-kernel variants against the product on C4 and C5 with the two-phase parity
# tests on each, and compute-sanitizer memcheck of the fuzz harness on the fastest-looking one (st4).
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
BRO_WORKLOADS=c4_highratio_w16,c5_stored_10k timeout 600 python tools/quick_perf.py "" lib_st3.so lib_st4.so lib_st5.so 2>&1 | tee gpurun_out/quick_variants.log
for v in lib_st3.so lib_st4.so lib_st5.so; do
  echo "== $v" | tee -a gpurun_out/pytest_gpu_variant.log
  BRO_B200_LIB=$PWD/brotli_rs_b200/lib/$v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "twophase or auto or side_by_side or size_hints or reservation" 2>&1 | tail -3 | tee -a gpurun_out/pytest_gpu_variant.log
done
BRO_B200_LIB=$PWD/brotli_rs_b200/lib/lib_st4.so timeout 600 compute-sanitizer --tool memcheck python tools/fuzz_gpu.py --count 300 2>&1 | tail -8 | tee gpurun_out/fuzz_memcheck_st4.log
