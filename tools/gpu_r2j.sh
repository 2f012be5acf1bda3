#!/bin/bash
# what bounds the corpus batch: every large corpus stream alone on both paths, and the source lines of the fused kernel on the slowest
mkdir -p gpurun_out
timeout 300 python tools/time_single.py 20000 2>&1 | tee gpurun_out/time_single.log
BRO_SINGLE_MODES=fused timeout 600 ncu --set full --clock-control none --import-source on -k regex:bro_decode_warp -c 1 -f -o gpurun_out/r02j_prof_mbreset \
    python tools/time_single.py 400000 metablock_reset > gpurun_out/r02j_prof_mbreset.log 2>&1
ls -la gpurun_out/r02j_prof_mbreset.ncu-rep
