"""Kernel-only timing sweep over the benchmark decks (scratch tool; prints one JSON line per case).

    python tools/perf_sweep.py [deck:nphoton[:accum] ...]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcxcl_b200 import benchmarks, engine, hostcfg

cases = sys.argv[1:] or ["cube60:1e7", "cube60b:1e7", "cube60b:1e8", "skinvessel:1e7"]
for c in cases:
    parts = c.split(":")
    name, nph = parts[0], float(parts[1])
    cfg = benchmarks.get(name, nph)
    if len(parts) > 2:
        cfg["accum"] = parts[2]
    for kv in parts[3:]:
        k, v = kv.split("=")
        cfg[k] = type(cfg.get(k, 0))(float(v)) if k in cfg else float(v)
    p = hostcfg.prepare(cfg)
    with engine.Simulation(p) as sim:
        ts = []
        for it in range(3):
            sim.reset()
            sim.launch()
            ts.append(sim.kernel_ms())
        r = sim.fetch()
        print(json.dumps(dict(case=c, kernel=sim.kernel_name, ms=[round(t, 3) for t in ts], pms=round(nph / min(ts), 1),
                              absorbed=round(r["absorbed"], 5), detected=r["detected"], nthread=r["nthread"])), flush=True)
