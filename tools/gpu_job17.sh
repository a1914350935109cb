#!/bin/bash
# boundary codes in the common kernels: sweep (regression check of the decks without codes), a deck WITH codes in both kernels, parity tests
mkdir -p gpurun_out
CASES="cube60:1e8 cube60b:1e8 skinvessel:1e8 colin27:3e7 digimouse:3e7" bash tools/gpu_job12.sh
python - <<'PY' > gpurun_out/r2_bc_common_vs_generic.log 2>&1
import os, sys, json
sys.path.insert(0, os.getcwd())
from mcxcl_b200 import benchmarks, engine, hostcfg
for name, bc, refl in (("cube60 -b 0 -B aarraa", "aarraa", 0), ("cube60 --bc cccccc", "cccccc", 1), ("cube60 --bc ______111111", "______111111", 0)):
    for env in ("", "1"):
        if env: os.environ["MCXB_BC_GENERIC"] = "1"
        else: os.environ.pop("MCXB_BC_GENERIC", None)
        n = 1e6 if "c" in bc[:6] else 3e7
        p = hostcfg.prepare(dict(benchmarks.get("cube60", n), bc=bc, isreflect=refl))
        with engine.Simulation(p) as sim:
            ms = []
            for _ in range(3):
                sim.reset(); sim.launch(); ms.append(sim.kernel_ms())
            r = sim.fetch()
            print(json.dumps(dict(case=name, kernel=sim.kernel_name, ms=round(min(ms), 2), absorbed=round(r["absorbed"], 5), detected=r["detected"])))
PY
cat gpurun_out/r2_bc_common_vs_generic.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py tests/test_gpu_exact.py -m gpu -q 2>&1 | tail -4
