#!/bin/bash
# one GPU session: tests, smoke, perf sweep, bench, ncu launch list + full capture of the photon kernel (results in gpurun_out/)
mkdir -p gpurun_out
ls /usr/lib/x86_64-linux-gnu/libnvidia-opencl* /etc/OpenCL/vendors 2>&1 | tee gpurun_out/opencl_probe.log
nproc | tee gpurun_out/nproc.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
python tools/perf_sweep.py cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:1e8 digimouse:1e8 digimouse_tg:1e8 cube60b:1e8:f64:issaveseed=1 cube60b:1e8:f64:tstep=5e-10 2>&1 | tee gpurun_out/sweep.log
python bench.py 2>&1 | tail -3 | tee gpurun_out/bench.log
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --photons 1e7 --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:photon_kernel -s 3 -c 1 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --photons 1e8 --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
ls -la gpurun_out
