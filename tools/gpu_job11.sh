#!/bin/bash
# 2 x B200: the C-ABI multi-GPU call with both exchanges, timings of the relinked CLI, and the new CLI tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cli.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest_multi2.log; cat gpurun_out/r2_pytest_multi2.log | tail -12
for ex in peer nccl; do
  ( cd /tmp && MCXB_TIMING=1 MCXB_MULTI_EXCHANGE=$ex timeout 600 /root/repo/integration/_build/mcxcl --bench colin27 -n 1e8 -G 11 -H 10000000 -S 0 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -E "mcxb multi|kernel complete|transfer complete|speed|absorbed" | sed "s/^/[$ex] /" ) >> gpurun_out/r2_cli_exchange.log 2>&1
done
cat gpurun_out/r2_cli_exchange.log
