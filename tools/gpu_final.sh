#!/bin/bash
# Short end-of-round GPU pass (gpurun -- 'bash tools/gpu_final.sh'): parity tests, the headline bench line with e2e and
# cpu_baseline, the reference arm, the ncu launch list and one --set full capture of the copy kernel, smoke.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c4_reference.json 2> gpurun_out/bench_c4_reference.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bro_ -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"bro_copy_kernel" -s 1 -c 1 -f -o gpurun_out/prof_c4_copy_r01c \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_c4_copy_r01c.log 2>&1
timeout 100 python bench.py --workload c5_stored_10k --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_c5_stored_10k.json 2> gpurun_out/bench_c5_stored_10k.err
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/smoke.log; cat gpurun_out/bench_c4.json gpurun_out/bench_c4_reference.json gpurun_out/bench_c5_stored_10k.json | cut -c1-900
