#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2_pytest_gpu.log
tail -5 gpurun_out/r2_pytest_gpu.log
: > gpurun_out/r2_sweep_aux.log
for v in default aux0; do
  if [ $v = default ]; then unset MCXB200_LIB; else export MCXB200_LIB=$PWD/mcxcl_b200/build/variants/$v/libmcxb200.so; fi
  echo "== $v" >> gpurun_out/r2_sweep_aux.log
  timeout 600 python tools/perf_sweep.py cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:3e7 digimouse:3e7 >> gpurun_out/r2_sweep_aux.log 2>&1
done
unset MCXB200_LIB
cut -c1-200 gpurun_out/r2_sweep_aux.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:photon_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_cube60b python tools/ncu_one.py cube60b 1e8 > gpurun_out/r2_ncu_cube60b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:photon_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_colin27 python tools/ncu_one.py colin27 3e7 > gpurun_out/r2_ncu_colin27.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:photon_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_skinvessel python tools/ncu_one.py skinvessel 3e7 > gpurun_out/r2_ncu_skinvessel.log 2>&1
tail -2 gpurun_out/r2_ncu_cube60b.log gpurun_out/r2_ncu_colin27.log
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['kernel_ms'], d['parity_check']['ok'], d['extra']['colin27']['value'], d['extra']['colin27']['e2e'])
PY
bash tools/deposit_variants.sh > /dev/null 2>&1
tail -70 gpurun_out/r2_deposit_variants.log
