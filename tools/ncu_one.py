"""One simulation for an ncu capture: python tools/ncu_one.py <deck> <photons> (1 warm-up launch + 1 profiled launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcxcl_b200 import benchmarks, engine, hostcfg

p = hostcfg.prepare(benchmarks.get(sys.argv[1], float(sys.argv[2])))
with engine.Simulation(p) as sim:
    for _ in range(2):
        sim.reset()
        sim.launch()
        print(sim.kernel_name, sim.kernel_ms())
