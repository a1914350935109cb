#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ext.py -m gpu -q > gpurun_out/r2_pytest_ext_full.log 2>&1
grep -E "^E  |^tests/|Error|passed|failed" gpurun_out/r2_pytest_ext_full.log | head -80 > gpurun_out/r2_pytest_ext.log; cat gpurun_out/r2_pytest_ext.log; tail -3 gpurun_out/r2_pytest_ext_full.log
