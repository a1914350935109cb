#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6; do timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ext.py tests/test_gpu_svmc.py -m gpu -q -p no:cacheprovider 2>&1 | tail -1; done
