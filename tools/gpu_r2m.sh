#!/bin/bash
# the copy kernel's two shapes on a rank's share of the headline batch under strong scaling (gpurun -- 'bash tools/gpu_r2m.sh')
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k twophase 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_quick.log
for n in 6250 12500 25000 50000 100000; do
  for sh in 0 1 auto; do
    if [ $sh = auto ]; then unset BRO_B200_COPY_SHAPE; else export BRO_B200_COPY_SHAPE=$sh; fi
    python bench.py --streams $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-write-roof --no-extra-workloads 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernels']
print('streams $n shape $sh: step %.3f ms  parse %.3f  copy %.3f' % (j['ms_per_step'], k['bro_parse_kernel']['ms'], k['bro_copy_kernel']['ms']))" | tee -a gpurun_out/copy_shapes.log
  done
done
