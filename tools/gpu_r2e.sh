#!/bin/bash
# bench.py check after the sampler / ceiling change: the N-rank line (clocks of all ranks, NUMA note, ceiling per trial); run under gpurun --gpus N
N=${1:-2}
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2e_bench_$N.json 2> gpurun_out/r2e_bench_$N.err
tail -3 gpurun_out/r2e_bench_$N.err
python - <<PY
import json
j=json.loads(open('gpurun_out/r2e_bench_$N.json').read().strip().splitlines()[-1])
print({k:j[k] for k in ['n_gpus','value','ms_per_step','clocks']})
print('strong', j.get('strong'))
print('e2e', j.get('e2e'))
PY
lscpu | grep -i 'numa\|socket\|^CPU(s)' | head; nvidia-smi topo -m | head -14
