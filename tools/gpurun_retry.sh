#!/bin/bash
# usage: tools/gpurun_retry.sh [--gpus N] <timeout> <script>   -- retries while the pod has no free slot (nothing is charged for those)
gp=""
if [ "$1" = "--gpus" ]; then gp="--gpus $2"; shift 2; fi
for attempt in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun $gp --timeout $1 -- "bash $2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"
  exit 0
done
echo "$out"; exit 3
