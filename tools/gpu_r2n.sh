#!/bin/bash
# the parse kernel's block size against the batch size (gpurun -- 'bash tools/gpu_r2n.sh'): the policy (auto) and forced 12 / 11 / 10 warps per SM
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 | tee gpurun_out/pytest_gpu_quick.log
for n in ${BRO_SIZES:-70000 100000 160000}; do
  for w in auto 12 11 10; do
    if [ $w = auto ]; then unset BRO_B200_PARSE_WARPS; else export BRO_B200_PARSE_WARPS=$w; fi
    python bench.py --streams $n --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --no-write-roof --no-extra-workloads 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); k=j['roofline']['kernels']
print('streams $n warps $w: step %.3f ms  parse %.3f  copy %.3f' % (j['ms_per_step'], k['bro_parse_kernel']['ms'], k['bro_copy_kernel']['ms']))" | tee -a gpurun_out/parse_warps2.log
  done
done
