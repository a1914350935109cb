"""Where the end-to-end time of one host-buffer call goes (scratch tool): create / reset / kernel / fetch / destroy."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mcxcl_b200 import benchmarks, engine, hostcfg

name = sys.argv[1] if len(sys.argv) > 1 else "cube60b"
nph = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8
p = hostcfg.prepare(benchmarks.get(name, nph))
engine.run_prepared(hostcfg.prepare(benchmarks.get(name, 1e5)))      # context + module load
for it in range(3):
    t = [time.perf_counter()]
    sim = engine.Simulation(p)
    t.append(time.perf_counter())
    sim.reset()
    t.append(time.perf_counter())
    sim.launch()
    ms = sim.kernel_ms()
    t.append(time.perf_counter())
    r = sim.fetch()
    t.append(time.perf_counter())
    sim.close()
    t.append(time.perf_counter())
    t0 = time.perf_counter()
    r2 = engine.run_prepared(p)
    t1 = time.perf_counter()
    d = np.diff(t) * 1e3
    print("%s %g: create %.2f reset %.2f kernel %.2f (event %.2f) fetch %.2f destroy %.2f | total %.2f | one-shot %.2f ms"
          % (name, nph, d[0], d[1], d[2], ms, d[3], d[4], (t[-1] - t[0]) * 1e3, (t1 - t0) * 1e3), flush=True)
