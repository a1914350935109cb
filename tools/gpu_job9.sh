#!/bin/bash
# extended physics (polarised, RF, adjoint) on the GPU, then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ext.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2_pytest_ext.log; cat gpurun_out/r2_pytest_ext.log | tail -30
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu2.log; tail -8 gpurun_out/r2_pytest_gpu2.log
