#!/bin/bash
# multi-GPU session (gpurun --gpus N): multi-GPU tests, then the bench at N GPUs for cube60b (default) and colin27
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/scale_gpus.log
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus $N --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_n$N.json
$TR bench.py --gpus $N --steps 3 --warmup 3 --workload colin27 --photons 1.25e8 2>&1 | tail -1 | tee gpurun_out/bench_colin27_n$N.json
python bench.py --gpus 1 --steps 3 --warmup 3 --workload colin27 --photons 1.25e8 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_colin27_n1.json
