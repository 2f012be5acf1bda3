#!/bin/bash
# source lines of the fused kernel on alice29 alone (C1) and of the parse kernel on c6 (immediate mode)
mkdir -p gpurun_out
BRO_SINGLE_MODES=fused timeout 600 ncu --set full --clock-control none --import-source on -k regex:bro_decode_warp -c 1 -f -o gpurun_out/r02k_prof_alice \
    python tools/time_single.py 50000 alice29 > gpurun_out/r02k_prof_alice.log 2>&1
ls -la gpurun_out/r02k_prof_alice.ncu-rep
