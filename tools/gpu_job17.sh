#!/bin/bash
# energy / length outputs in the BCODES common kernels vs the generic kernels, then the parity tests that cover output types
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2_ot_common_vs_generic.log 2>&1
import os, sys, json
sys.path.insert(0, os.getcwd())
import numpy as np
from mcxcl_b200 import benchmarks, engine, hostcfg
for ot in ("energy", "length"):
    res = {}
    for env in ("", "1"):
        if env: os.environ["MCXB_BC_GENERIC"] = "1"
        else: os.environ.pop("MCXB_BC_GENERIC", None)
        p = hostcfg.prepare(dict(benchmarks.get("cube60b", 3e7), outputtype=ot, issavedet=0))
        with engine.Simulation(p) as sim:
            ms = []
            for _ in range(3):
                sim.reset(); sim.launch(); ms.append(sim.kernel_ms())
            r = sim.fetch()
            res[env] = r["field"].astype(np.float64).sum()
            print(json.dumps(dict(case="cube60b -O " + ot, kernel=sim.kernel_name, ms=round(min(ms), 2), absorbed=round(r["absorbed"], 5), fieldsum=res[env])))
    print("  ratio of the field sums (common / generic): %.5f" % (res[""] / res["1"]))
PY
cat gpurun_out/r2_ot_common_vs_generic.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -3
