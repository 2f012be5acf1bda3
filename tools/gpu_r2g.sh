#!/bin/bash
# Round 2, second half (gpurun -- 'bash tools/gpu_r2g.sh'): variants of the copy kernel (sliding window in shared memory, pinned
# %tid values) and of the parse kernel (pinned %tid values) on the headline batch and C5 / C5b, the two-phase parity tests on the
# window variants, and one --set full capture of the parse kernel (source lines of the whole kernel).
mkdir -p gpurun_out
BRO_WORKLOADS=c4_highratio_w16,c5_stored_10k,c5b_literals_10k timeout 500 python tools/quick_perf.py "" lib_nopin.so lib_cpin.so lib_win.so lib_winpin.so 2>&1 | tee gpurun_out/quick_variants.log
for v in lib_win.so lib_winpin.so; do
  echo "== $v" | tee -a gpurun_out/pytest_gpu_variant.log
  BRO_B200_LIB=$PWD/brotli_rs_b200/lib/$v timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "twophase" 2>&1 | tail -3 | tee -a gpurun_out/pytest_gpu_variant.log
done
bash tools/gpu_prof2.sh bro_parse_kernel c4_highratio_w16 r02g_prof_c4_parse "" 0
