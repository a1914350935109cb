#!/bin/bash
# last check of the round: the driver's own sequence (GPU suite with -x, smoke, both bench arms)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-200
timeout 900 python bench.py > gpurun_out/r2_bench_last.json 2>/dev/null; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_last.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['parity_check']['ok'], {k:(round(v['value']), v['parity_check']['ok']) for k,v in d['extra'].items()})
PY
