#!/bin/bash
# 1 -> N GPU scaling of the headline bench, launched exactly like the driver does (torchrun for N > 1).
# usage: scale_round.sh "<list of N>" [weak|strong]   e.g.  scale_round.sh "1 2"   or   scale_round.sh "4 8" strong
mkdir -p gpurun_out
NS=${1:-"1 2"}
SC=${2:-weak}
nvidia-smi -L | head -8
NMAX=1
for n in $NS; do
  NMAX=$n
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --scaling $SC > gpurun_out/scale_${SC}_n1.json 2> gpurun_out/scale_${SC}_n1.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
        bench.py --gpus $n --steps 10 --warmup 3 --scaling $SC > gpurun_out/scale_${SC}_n$n.json 2> gpurun_out/scale_${SC}_n$n.err
  fi
done
if [ $NMAX -gt 1 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --impl reference --gpus $NMAX --steps 3 --warmup 1 > gpurun_out/scale_ref_n$NMAX.json 2> gpurun_out/scale_ref_n$NMAX.err
fi
tail -n 2 gpurun_out/scale_${SC}_*.err
cat gpurun_out/scale_${SC}_*.json gpurun_out/scale_ref_*.json | cut -c1-300
