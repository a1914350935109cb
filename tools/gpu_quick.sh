mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
python tools/perf_sweep.py cube60:1e8 cube60b:1e8 skinvessel:1e7 colin27:1e7 digimouse:1e7 2>&1 | tee gpurun_out/sweep.log
python tools/e2e_breakdown.py 2>&1 | tee gpurun_out/e2e_breakdown.log
