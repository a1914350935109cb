#!/bin/bash
# 8 x B200: the N=8 bench line (photon shards + NCCL), the CPU arm under torchrun, and the relinked CLI with -G 11111111
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_gpus8.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench8 rc $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n8.json') if l.startswith('{')][-1])
print("N=8 value", d['value'], "e2e", d['e2e']['value'], "parity", d['parity_check']['ok'], d['parity_check'].get('detected_whole_job'), "colin27", d['extra']['colin27']['value'], d['extra']['colin27']['e2e'], d['extra']['colin27']['parity_check']['ok'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_n8.json 2> gpurun_out/r2_bench_ref_n8.err
grep -o '"value": [0-9.]*, "unit": "photons/ms", "n_gpus": 8\|"threads": [0-9]*\|"omp_num_threads_env": [^,]*' gpurun_out/r2_bench_ref_n8.json | head -4
( cd /tmp && timeout 600 /root/repo/integration/_build/mcxcl --bench colin27 -n 1e9 -G 11111111 -H 10000000 -S 0 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -E "device|kernel complete|transfer complete|detected|simulated|speed|absorbed|NCCL" ) > gpurun_out/r2_cli_multi.log 2>&1
( cd /tmp && timeout 600 /root/repo/integration/_build/mcxcl --bench colin27 -n 1e9 -G 1 -H 10000000 -S 0 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -E "device|kernel complete|transfer complete|detected|simulated|speed|absorbed" ) >> gpurun_out/r2_cli_multi.log 2>&1
cat gpurun_out/r2_cli_multi.log
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
