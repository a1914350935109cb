#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pmcxcl.py -m gpu -q 2>&1 | grep -E "^E|^tests/|Error|passed|failed" | head -60 > gpurun_out/r2_pytest_ext.log; cat gpurun_out/r2_pytest_ext.log
