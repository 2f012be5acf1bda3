#!/bin/bash
# ncu evidence for the decode kernel (B200_PROFILING.md recipe): launch list of the default bench command and one
# --set full capture of the kernel on a reduced batch (ncu replays each launch ~40 times).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bro_decode_kernel -s 2 -c 1 -o gpurun_out/prof_c4 \
    python bench.py --steps 2 --warmup 1 --streams 20000 --no-e2e --no-cpu-baseline > gpurun_out/prof_c4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bro_decode_kernel -s 2 -c 1 -o gpurun_out/prof_c2 \
    python bench.py --workload c2_quickfox_x10k --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_c2.log 2>&1
ls -la gpurun_out
