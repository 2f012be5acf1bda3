#!/bin/bash
# usage: bash tools/gpu_prof2.sh <kernel regex> <workload> <out name> <lib or ""> [skip]
mkdir -p gpurun_out
K=$1; W=$2; O=$3; L=$4; S=${5:-1}
if [ -n "$L" ]; then export BRO_B200_LIB=$PWD/brotli_rs_b200/lib/$L; fi
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c 1 -f -o gpurun_out/$O \
    python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/$O.log 2>&1
ls -la gpurun_out/$O.ncu-rep
