#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -m gpu -q -k "polarised or rf_frequency" > gpurun_out/r2_pytest_ext_full.log 2>&1
grep -E "^E  |^tests/|Error|passed|failed" gpurun_out/r2_pytest_ext_full.log | head -60; tail -3 gpurun_out/r2_pytest_ext_full.log
