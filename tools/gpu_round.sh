#!/bin/bash
# One GPU-box round: parity tests, headline bench, the other BASELINE configs (kernel-only), ncu launch list.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
for w in c2_quickfox_x10k c3_corpus_x1000 c5_stored_10k c5b_literals_10k; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
tail -3 gpurun_out/*.err
cat gpurun_out/pytest_gpu.log gpurun_out/bench_*.json
