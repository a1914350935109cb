#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus2.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log
tail -6 gpurun_out/r2_pytest_gpu.log
timeout 600 python tools/perf_sweep.py cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:3e7 digimouse:3e7 > gpurun_out/r2_sweep_q.log 2>&1
cut -c1-200 gpurun_out/r2_sweep_q.log
( cd /tmp && timeout 600 /root/repo/integration/_build/mcxcl --bench colin27 -n 1e8 -G 11 -H 1e7 -S 0 2>&1 | tail -12 ) > gpurun_out/r2_cli_g11.log 2>&1
( cd /tmp && timeout 600 /root/repo/integration/_build/mcxcl --bench colin27 -n 1e8 -G 1 -H 1e7 -S 0 2>&1 | tail -8 ) > gpurun_out/r2_cli_g1.log 2>&1
cat gpurun_out/r2_cli_g11.log gpurun_out/r2_cli_g1.log
