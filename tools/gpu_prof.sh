#!/bin/bash
# one --set full capture (with source) of the kernels matching $1 on workload $2 (default: parse kernel, C4)
# usage: bash tools/gpu_prof.sh <kernel regex> <workload> <out name> [skip launches]
mkdir -p gpurun_out
K=${1:-bro_parse_kernel}; W=${2:-c4_highratio_w16}; O=${3:-prof}; S=${4:-1}
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -f -o gpurun_out/$O \
    python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/$O.log 2>&1
tail -2 gpurun_out/$O.log | cut -c1-300
ls -la gpurun_out/$O.ncu-rep
