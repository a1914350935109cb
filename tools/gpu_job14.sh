#!/bin/bash
# final round-2 pass on one B200: whole GPU suite, smoke, bench (both arms), launch list of the bench command, digimouse counters
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_gpu_final.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu_final.log
tail -6 gpurun_out/r2_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_final.json 2> gpurun_out/r2_bench_ref_final.err; echo "ref rc $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2_ncu_launch_bench.log 2>&1; echo "ncu rc $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:photon_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_digimouse_v10 python tools/ncu_one.py digimouse 3e7 > gpurun_out/r2_ncu_digimouse_v10.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['kernel_ms'], d['parity_check']['ok'], d['roofline']['frac'], {k:(v['value'], v['e2e'], v['parity_check']['ok']) for k,v in d['extra'].items()}, d['cpu_baseline']['value'])
PY
