#!/bin/bash
# One GPU-box round: parity tests, smoke, headline bench (with e2e + cpu baseline), the other BASELINE configs
# (kernel-only), the reference arm, and the ncu evidence (launch list + one --set full capture per workload).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c4_reference.json 2> gpurun_out/bench_c4_reference.err
for w in c2_quickfox_x10k c3_corpus_x1000 c5_stored_10k c5b_literals_10k; do
  python bench.py --workload $w --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
for w in c4_highratio_w16:20000 c2_quickfox_x10k:10000 c5_stored_10k:100000; do
  ncu --set full --clock-control none --import-source on -k regex:bro_decode -s 2 -c 1 -o gpurun_out/prof_${w%%:*} \
      python bench.py --workload ${w%%:*} --streams ${w##*:} --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_${w%%:*}.log 2>&1
done
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cat gpurun_out/bench_*.json | cut -c1-600
