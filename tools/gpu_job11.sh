#!/bin/bash
# N x B200: timings of the relinked CLI with the peer-memory exchange (and NCCL for comparison)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); MASK=$(printf '1%.0s' $(seq 1 $N))
: > gpurun_out/r2_cli_exchange.log
for ex in peer nccl; do
  ( cd /tmp && MCXB_TIMING=1 MCXB_MULTI_EXCHANGE=$ex timeout 600 /root/repo/integration/_build/mcxcl --bench colin27 -n ${1:-1e8} -G $MASK -H 10000000 -S 0 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -E "mcxb multi|kernel complete|transfer complete|speed|absorbed" | sed "s/^/[$ex x$N] /" ) >> gpurun_out/r2_cli_exchange.log 2>&1
done
cat gpurun_out/r2_cli_exchange.log
