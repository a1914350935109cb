"""Scratch: what the ~20 ms of fixed cost of a 1e8-photon launch is made of (clock ramp? L2 flush? nvidia-smi? torch?)."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcxcl_b200 import benchmarks, engine, hostcfg


def run(sim, n=3, before=None):
    ts = []
    for _ in range(n):
        sim.reset()
        if before:
            before()
        sim.launch()
        ts.append(round(sim.kernel_ms(), 2))
    return ts


print("== linearity")
for nph in (1e7, 2.5e7, 5e7, 1e8, 2e8, 4e8):
    with engine.Simulation(hostcfg.prepare(benchmarks.get("cube60b", nph))) as sim:
        ts = run(sim, 2)
        print("N=%g ms=%s  ms per 1e8=%.2f" % (nph, ts, min(ts) * 1e8 / nph), flush=True)

p = hostcfg.prepare(benchmarks.get("cube60b", 1e8))
with engine.Simulation(p) as sim:
    print("plain            ", run(sim), flush=True)
    time.sleep(1.0)
    print("after 1 s idle   ", run(sim, 1), flush=True)
    import torch
    torch.cuda.set_device(0)
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    print("torch imported   ", run(sim), flush=True)
    print("flush before     ", run(sim, before=lambda: flush.zero_()), flush=True)

    def spin():
        for _ in range(60):
            torch.matmul(a, a)
    print("matmul 60x before", run(sim, before=spin), flush=True)
    proc = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "200"],
                            stdout=subprocess.PIPE, text=True)
    time.sleep(0.5)
    print("nvidia-smi -lms  ", run(sim), flush=True)
    proc.terminate()
    print("clocks seen      ", proc.stdout.read().split("\n")[:12])
    print("plain again      ", run(sim), flush=True)
