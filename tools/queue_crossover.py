"""Where does the scattering queue stop paying?  cube60b with the scattering coefficient of its medium swept, kernels
with (MCXB_SCATTER_QUEUE=1) and without (=0) the queue.  Scratch tool: one JSON line per case."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcxcl_b200 import benchmarks, engine, hostcfg

nph = float(sys.argv[1]) if len(sys.argv) > 1 else 3e7
for mus in (0.25, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 4.0, 6.0, 10.0):
    row = dict(mus=mus)
    for q in ("0", "1"):
        os.environ["MCXB_SCATTER_QUEUE"] = q
        cfg = benchmarks.get("cube60b", nph)
        cfg["prop"] = [[0, 0, 1, 1], [0.005, mus, 0.01 if mus <= 1 else 0.9, 1.37], [0.002, 5.0, 0.9, 1.0]]
        p = hostcfg.prepare(cfg)
        with engine.Simulation(p) as sim:
            ts = []
            for _ in range(3):
                sim.reset()
                sim.launch()
                ts.append(sim.kernel_ms())
            r = sim.fetch()
            row["q" + q] = round(min(ts), 2)
            row["k" + q] = sim.kernel_name.split("/")[-1]
            row["abs" + q] = round(r["absorbed"], 5)
    row["gain"] = round(row["q0"] / row["q1"], 4)
    print(json.dumps(row), flush=True)
