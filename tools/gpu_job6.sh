#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
: > gpurun_out/r2_sweep_u2.log
for v in default u2; do
  if [ $v = default ]; then unset MCXB200_LIB; else export MCXB200_LIB=$PWD/mcxcl_b200/build/variants/$v/libmcxb200.so; fi
  echo "== $v" >> gpurun_out/r2_sweep_u2.log
  timeout 600 python tools/perf_sweep.py cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:3e7 digimouse:3e7 >> gpurun_out/r2_sweep_u2.log 2>&1
done
unset MCXB200_LIB
cut -c1-200 gpurun_out/r2_sweep_u2.log
