#!/bin/bash
# retry a gpurun call until the pod has a slot (exit code 3 = transient); usage: tools/gpurun_retry.sh <timeout> '<command>'
T=$1; shift
for i in $(seq 1 40); do
  gpurun --timeout $T -- "$@" > gpurun_out/.retry_last.log 2>&1
  rc=$?
  if ! grep -q "status=transient" gpurun_out/.retry_last.log; then cat gpurun_out/.retry_last.log | tail -${TAILN:-15}; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
