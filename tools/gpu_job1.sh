#!/bin/bash
# round-2 first contact: full GPU test suite, default bench line, reference arm
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus.log 2>&1; nproc >> gpurun_out/r2_gpus.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc $?"
tail -c 1500 gpurun_out/r2_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>&1
tail -c 600 gpurun_out/r2_bench_ref.json
