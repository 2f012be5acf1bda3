#!/bin/bash
# One GPU-box round: parity tests, smoke, headline bench (with e2e + cpu baseline), the other BASELINE configs
# (kernel-only), the reference arm, and the ncu evidence (launch list + --set full captures of every kernel of the path).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 1500 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_c4_reference.json 2> gpurun_out/bench_c4_reference.err
timeout 900 python bench.py --mode warp --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c4_warp.json 2> gpurun_out/bench_c4_warp.err
for w in c2_quickfox_x10k c3_corpus_x1000 c5_stored_10k c5b_literals_10k; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bro_ -c 60 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# one full capture per kernel of the headline step (launches 1.. of the warm-up call: parse, copy, fused retry)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bro_parse_kernel|bro_copy_kernel" -s 2 -c 2 -f -o gpurun_out/prof_c4_highratio_w16 \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_c4_highratio_w16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bro_decode_warp -s 1 -c 1 -f -o gpurun_out/prof_c2_quickfox_x10k \
    python bench.py --workload c2_quickfox_x10k --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_c2_quickfox_x10k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bro_parse_kernel|bro_copy_kernel" -s 2 -c 2 -f -o gpurun_out/prof_c5b_literals_10k \
    python bench.py --workload c5b_literals_10k --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/prof_c5b_literals_10k.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cat gpurun_out/bench_*.json | cut -c1-700
