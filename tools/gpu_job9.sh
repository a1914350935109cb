#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "boundary_codes_and_energy" > gpurun_out/r2_pytest_ext_full.log 2>&1
grep -E "^E  |^tests/|Error|passed|failed" gpurun_out/r2_pytest_ext_full.log | head -40; tail -3 gpurun_out/r2_pytest_ext_full.log
