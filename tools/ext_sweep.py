"""Kernel-window timings of the extended modes (EXT kernels): python tools/ext_sweep.py [photons]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from mcxcl_b200 import benchmarks, engine, hostcfg
from test_gpu_ext import rayleigh
from test_gpu_svmc import tilted_slab

n = float(sys.argv[1]) if len(sys.argv) > 1 else 1e7
base = benchmarks.get("cube60b", n)
cases = {
    "cube60b (common kernel, for scale)": dict(base, issavedet=0),
    "cube60b generic kernel (gscatter set)": dict(base, issavedet=0, gscatter=999999999),
    "cube60b polarised (Rayleigh matrix)": dict(base, prop=[[0, 0, 1, 1], [0.005, 1.0, 0.0, 1.37]], smatrix=rayleigh(1), srciquv=[1, 1, 0, 0], savedetflag="dpi"),
    "cube60b RF forward 100 MHz": dict(base, issavedet=0, omega=2 * np.pi * 1e8),
    "split-voxel slab 40^3": dict(vol=tilted_slab(), prop=[[0, 0, 1, 1], [0.02, 1.0, 0.8, 1.37], [0.005, 2.0, 0.9, 1.55]], nphoton=n, srcpos=[20, 20, 0],
                                  srcdir=[0, 0, 1], issrcfrom0=1, tstart=0, tend=5e-9, tstep=5e-9, isreflect=1, seed=12345, issavedet=0),
    "cube60 adjoint (1 source + 1 detector disk)": dict(benchmarks.get("cube60", n), issavedet=0, srcpos=[[30, 30, 1, 1], [30, 42, 1, 1]],
                                                       srcdir=[[0, 0, 1, 0], [0, 0, 1, 0]], srcparam1=[[0, 0, 0, 0], [4, 0, 0, 0]], detpos=[[30, 42, 1, 4]],
                                                       srcid=-1, outputtype="adjoint"),
}
for name, cfg in cases.items():
    p = hostcfg.prepare(cfg)
    with engine.Simulation(p) as sim:
        ms = []
        for _ in range(3):
            sim.reset()
            sim.launch()
            ms.append(sim.kernel_ms())
        r = sim.fetch()
        print(json.dumps(dict(case=name, kernel=sim.kernel_name, photons=n, ms=[round(x, 2) for x in ms], pms=round(n / min(ms), 1), absorbed=round(r["absorbed"], 5))))
