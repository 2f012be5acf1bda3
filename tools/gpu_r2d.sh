#!/bin/bash
# multi-GPU check (run under gpurun --gpus N): bro_mg_* on every device of the box, then the N-rank bench with the strong and exchange records
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "mg_decode" 2>&1 | tail -3 | tee gpurun_out/r2d_mg_pytest_$N.log
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2d_bench_$N.json 2> gpurun_out/r2d_bench_$N.err
tail -3 gpurun_out/r2d_bench_$N.err
python - <<PY
import json
j=json.loads(open('gpurun_out/r2d_bench_$N.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ['n_gpus','value','ms_per_step']})
print('strong', j.get('strong'))
print('exchange', j.get('exchange'))
print('e2e', j.get('e2e'))
PY
