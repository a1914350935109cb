"""Scratch: phases of repeated one-shot calls (MCXB_TIMING=1), with and without a torch context in the process."""
import os
import sys
import time

os.environ["MCXB_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcxcl_b200 import benchmarks, engine, hostcfg

p = hostcfg.prepare(benchmarks.get("cube60b", 1e8))
for phase in ("plain", "torch"):
    if phase == "torch":
        import torch
        flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device="cuda:0")
        flush.zero_()
        torch.cuda.synchronize()
    for i in range(5):
        t0 = time.perf_counter()
        r = engine.run_prepared(p)
        print("%s call %d: wall %.2f ms" % (phase, i, (time.perf_counter() - t0) * 1e3), flush=True)
