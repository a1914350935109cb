#!/bin/bash
# validation of the shipped build (marcher inlined in the common / extended kernels): extended-mode timings, whole GPU suite, default bench line
mkdir -p gpurun_out
python tools/ext_sweep.py 1e7 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  %-45s %8.2f ms' % (d['case'], min(d['ms'])))
"
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo bench rc $?
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline'], d.get('parity'))
for k,v in d.get('extra',{}).items(): print(k, v.get('value'), v.get('e2e',{}).get('value') if isinstance(v.get('e2e'),dict) else v.get('e2e'))
PY
