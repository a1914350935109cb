#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log
tail -4 gpurun_out/r2_pytest_gpu.log
: > gpurun_out/r2_sweep_launch.log
echo "== default (plain launch + unroll 2)" >> gpurun_out/r2_sweep_launch.log
timeout 600 python tools/perf_sweep.py cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:3e7 digimouse:3e7 digimouse_tg:3e7 >> gpurun_out/r2_sweep_launch.log 2>&1
echo "== MCXB_NO_PLAINLAUNCH=1" >> gpurun_out/r2_sweep_launch.log
MCXB_NO_PLAINLAUNCH=1 timeout 600 python tools/perf_sweep.py cube60:1e8 cube60b:1e8 colin27:3e7 >> gpurun_out/r2_sweep_launch.log 2>&1
echo "== MCXB_SCATTER_QUEUE=0" >> gpurun_out/r2_sweep_launch.log
MCXB_SCATTER_QUEUE=0 timeout 600 python tools/perf_sweep.py cube60:1e8 cube60b:1e8 skinvessel:1e8 >> gpurun_out/r2_sweep_launch.log 2>&1
cut -c1-200 gpurun_out/r2_sweep_launch.log
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['kernel_ms'], d['parity_check']['ok'], d['roofline']['frac'], d['extra']['colin27']['value'], d['extra']['colin27']['e2e'])
PY
