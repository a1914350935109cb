#!/bin/bash
# memcheck of the extended-physics kernels: the timing sweep and the parity tests themselves under compute-sanitizer
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/ext_sweep.py 3e6 > gpurun_out/r2_memcheck_ext.log 2>&1; echo "rc $?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned" gpurun_out/r2_memcheck_ext.log | head -20
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_svmc.py tests/test_gpu_ext.py tests/test_gpu_media.py -m gpu -q -x > gpurun_out/r2_memcheck_tests.log 2>&1; echo "rc $?"
grep -E "ERROR SUMMARY|Invalid|out of bounds|misaligned|passed|failed" gpurun_out/r2_memcheck_tests.log | head -20
