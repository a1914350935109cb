#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_svmc.py tests/test_gpu_ext.py tests/test_gpu_pmcxcl.py -m gpu -q > gpurun_out/r2_pytest_ext_full.log 2>&1
grep -E "^E  |^tests/|Error|passed|failed" gpurun_out/r2_pytest_ext_full.log | head -60; tail -3 gpurun_out/r2_pytest_ext_full.log
