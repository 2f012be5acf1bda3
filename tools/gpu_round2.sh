#!/bin/bash
# Round-2 GPU box round: parity tests, smoke, headline bench (with e2e + cpu baseline + extra workloads), the reference arm, and the
# ncu evidence (launch list + --set full captures of every kernel of the path).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1
timeout 1500 python bench.py > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_c4_reference.json 2> gpurun_out/r02_bench_c4_reference.err
timeout 900 python bench.py --mode warp --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_c4_warp.json 2> gpurun_out/r02_bench_c4_warp.err
for w in c1_alice29_single c6_text_q11_w16 c7_far_w22; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-write-roof > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bro_ -c 60 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-write-roof > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bro_parse_kernel|bro_copy_kernel" -s 2 -c 2 -f -o gpurun_out/r02_prof_c4 \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-write-roof > gpurun_out/r02_prof_c4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bro_decode_warp -s 1 -c 1 -f -o gpurun_out/r02_prof_c2 \
    python bench.py --workload c2_quickfox_x10k --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-write-roof > gpurun_out/r02_prof_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bro_parse_kernel" -s 1 -c 1 -f -o gpurun_out/r02_prof_c5b \
    python bench.py --workload c5b_literals_10k --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-write-roof > gpurun_out/r02_prof_c5b.log 2>&1
cat gpurun_out/r02_pytest_gpu.log gpurun_out/r02_smoke.log; cut -c1-400 gpurun_out/r02_bench_*.json
