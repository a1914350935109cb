#!/bin/bash
# bench with the three extra legs on N GPUs (N = what the box has); checks the threaded host loop of fetch on the way
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
if [ "$N" -gt 1 ]; then
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_tg_n$N.json 2> gpurun_out/r2_bench_tg_n$N.err; echo "bench rc $?"
else
  (time timeout 1200 python bench.py) > gpurun_out/r2_bench_tg_n1.json 2> gpurun_out/r2_bench_tg_n1.err; echo "bench rc $?"; tail -4 gpurun_out/r2_bench_tg_n1.err
fi
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_tg_n$N.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['parity_check']['ok'], {k:(round(v['value']), round(v['e2e']), v['parity_check']['ok'], v.get('issue_frac')) for k,v in d['extra'].items()})
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "digimouse or time" 2>&1 | tail -2
