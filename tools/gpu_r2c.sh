#!/bin/bash
# full GPU suite, default bench line, and small-batch (strong-scaling share) timings with and without fewer streams per parse warp
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r2c_pytest.log
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; tail -3 gpurun_out/r2c_bench.err; cut -c1-1500 gpurun_out/r2c_bench.json
for n in 12500 25000 50000; do
  for lanes in 0 32; do
    echo "== streams $n lanes $lanes"
    BRO_B200_PARSE_LANES=$lanes timeout 300 python bench.py --streams $n --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-write-roof --mode twophase 2>&1 | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('%.3f ms  %.0f GB/s' % (j['ms_per_step'], j['value']), {k.split('_kernel')[0]: round(v['ms'],3) for k,v in j['roofline']['kernels'].items()})"
  done
done 2>&1 | tee gpurun_out/r2c_small.log
