#!/bin/bash
# c6 (context-modelled text) through the two-phase path at several batch sizes: per-stream latency vs L2 residency of the tables
for n in 1500 5000 10000 20000 40000; do
  python bench.py --workload c6_text_q11_w16 --streams $n --mode twophase --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-write-roof 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1])
print($n, 'streams', round(j['ms_per_step'],2), 'ms', round(j['value'],1), 'GB/s', {k.replace('bro_','').replace('_kernel',''):round(v['ms'],2) for k,v in j['roofline']['kernels'].items()})
"
done
