#!/bin/bash
# A/B of kernel variants: perf sweep of the shipped build and of every variant library given on the command line
mkdir -p gpurun_out
echo "== shipped build" > gpurun_out/r2_sweep_ab.log
timeout 600 python tools/perf_sweep.py ${CASES:-cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:3e7 digimouse:3e7} >> gpurun_out/r2_sweep_ab.log 2>&1
for v in "$@"; do
  echo "== variant $v" >> gpurun_out/r2_sweep_ab.log
  MCXB200_LIB=$v timeout 600 python tools/perf_sweep.py ${CASES:-cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:3e7 digimouse:3e7} >> gpurun_out/r2_sweep_ab.log 2>&1
done
python - <<'PY'
import json
for l in open('gpurun_out/r2_sweep_ab.log'):
    if l.startswith('=='): print(l.strip())
    elif l.startswith('{'):
        d=json.loads(l); print("  %-16s %8.2f ms  %s" % (d['case'], min(d['ms']), d['kernel']))
PY
