#!/bin/bash
mkdir -p gpurun_out
python tools/ext_sweep.py 1e7 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('  %-45s %8.2f ms' % (d['case'], min(d['ms'])))
"
for i in 1 2 3; do timeout 900 python -m pytest tests/test_gpu_ext.py tests/test_gpu_pmcxcl.py tests/test_gpu_cli.py -m gpu -q -p no:cacheprovider -k "polar" 2>&1 | tail -1; done
