#!/bin/bash
# after the shared-window change of the queue-less kernels: sweep, parity tests, fresh ncu capture of colin27
mkdir -p gpurun_out
CASES="cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:3e7 digimouse:3e7 digimouse_tg:3e7" bash tools/gpu_job12.sh
timeout 1200 python -m pytest tests/test_gpu_exact.py tests/test_gpu_parity.py tests/test_replay.py -m gpu -q 2>&1 | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:photon_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_colin27_v10 python tools/ncu_one.py colin27 3e7 > gpurun_out/r2_ncu_colin27_v10.log 2>&1
tail -2 gpurun_out/r2_ncu_colin27_v10.log
