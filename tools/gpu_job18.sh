#!/bin/bash
# repeatability of the statistical GPU tests: the whole suite N times (dynamic photon scheduling makes every run a new realisation)
mkdir -p gpurun_out
: > gpurun_out/r2_pytest_repeat.log
for i in 1 2 3; do
  timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4 >> gpurun_out/r2_pytest_repeat.log
done
grep -E "passed|failed|FAILED" gpurun_out/r2_pytest_repeat.log
