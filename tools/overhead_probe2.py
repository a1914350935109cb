"""Scratch: does the order of allocations (torch flush buffer before/after the simulation) change the kernel time?"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcxcl_b200 import benchmarks, engine, hostcfg

mode = sys.argv[1]
p = hostcfg.prepare(benchmarks.get("cube60b", 1e8))
if mode == "cuda_first":
    # a plain 384 MB cudaMalloc through ctypes, no torch
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    ptr = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(ptr), ctypes.c_size_t(384 << 20)) == 0
if mode == "small_first":
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    ptr = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(ptr), ctypes.c_size_t(1 << 20)) == 0
if mode in ("torch_first", "torch_first_reseed", "torch_first_finalize", "torch_freed"):
    import torch
    torch.cuda.set_device(0)
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    flush.zero_()
    torch.cuda.synchronize()
    if mode == "torch_freed":
        del flush
        torch.cuda.empty_cache()
sim = engine.Simulation(p)
if mode == "torch_first_reseed":
    sim.reseed(p.c.seed, 0)
ts = []
for _ in range(3):
    sim.reset()
    sim.launch()
    if mode == "torch_first_finalize":
        sim.finalize()
    ts.append(round(sim.kernel_ms(), 2))
print(mode, ts, flush=True)
