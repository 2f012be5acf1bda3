#!/bin/bash
# Verification pass of the committed tree (gpurun -- 'bash tools/gpu_r2f.sh'): GPU parity suite, smoke, the default bench run
# exactly as the driver starts it (wall time recorded), the reference arm.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
( time timeout 600 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
( time timeout 300 python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; tail -4 gpurun_out/bench_default.err; cut -c1-1500 gpurun_out/bench_default.json; tail -4 gpurun_out/bench_reference.err; cut -c1-600 gpurun_out/bench_reference.json
