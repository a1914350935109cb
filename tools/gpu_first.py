"""First contact with the GPU: parity spot checks + first timings (scratch tool; results land in gpurun_out/)."""
import ctypes as C
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from mcxcl_b200 import abi, benchmarks, engine, hostcfg
from oracle import loader

out = {}
os.makedirs("gpurun_out", exist_ok=True)
lib = abi.load()
out["gpus"] = engine.gpuinfo()
print(out["gpus"])
ref = loader.ref()

# --- RNG
seeds = ref.seeds(1648335518, 4096)
want, wst = ref.rng(seeds, 64)
got = np.zeros_like(want)
gst = np.zeros_like(wst)
abi.check(lib.mcxb_test_rng(0, seeds.ctypes.data, 4096, 64, got.ctypes.data, gst.ctypes.data))
out["rng_bitexact"] = bool((got.view(np.uint32) == want.view(np.uint32)).all() and (gst == wst).all())
print("rng bitexact", out["rng_bitexact"])

# --- traversal
rs = np.random.RandomState(7)
n = 20000
p0 = np.zeros((n, 4), np.float32)
p0[:, :3] = rs.uniform(0, 60, (n, 3))
v0 = np.zeros((n, 4), np.float32)
d = rs.normal(size=(n, 3))
d /= np.linalg.norm(d, axis=1, keepdims=True)
v0[:, :3] = d
for musp in (1.0, 0.731, 9.398496241 * 0.005):
    want = ref.trace(p0, v0, 48, (60, 60, 60), musp)
    buf = (abi.TraceStep * (n * 48))()
    abi.check(lib.mcxb_test_trace(0, p0.ctypes.data, v0.ctypes.data, n, 48, 60, 60, 60, C.c_float(musp), C.addressof(buf)))
    got = np.frombuffer(buf, dtype=loader.TRACE_DTYPE).reshape(n, 48)
    same = all((got[f].view(np.uint32) if got[f].dtype == np.float32 else got[f]).tobytes() == (want[f].view(np.uint32) if want[f].dtype == np.float32 else want[f]).tobytes() for f in got.dtype.names)
    out["trace_bitexact_%g" % musp] = bool(same)
    print("trace bitexact musp", musp, same)

# --- physics
for name in ("cube60", "cube60b"):
    p = hostcfg.prepare(benchmarks.get(name, 1e6))
    r = engine.run_prepared(p)
    o = ref.run(hostcfg.prepare(benchmarks.get(name, 2e5)), 4096, hostthreads=0)
    out[name] = dict(absorbed=r["absorbed"], ref_absorbed=o["absorbed"], detected=r["detected"], ref_detected_2e5=o["detected"],
                     ms=r["runtime_ms"], pms=1e6 / r["runtime_ms"], energytot=r["energytot"], nthread=r["nthread"])
    print(name, out[name])

# --- timing sweep
for name, nph in (("cube60", 1e7), ("cube60b", 1e7), ("cube60b", 1e8), ("cube60", 1e8)):
    for accum in ("f64", "f32"):
        cfg = benchmarks.get(name, nph)
        cfg["accum"] = accum
        p = hostcfg.prepare(cfg)
        with engine.Simulation(p) as sim:
            ts = []
            for it in range(3):
                sim.reset()
                sim.launch()
                ts.append(sim.kernel_ms())
            r = sim.fetch()
        key = "%s_%g_%s" % (name, nph, accum)
        out[key] = dict(ms=ts, pms=nph / min(ts), absorbed=r["absorbed"], kernel=sim.kernel_name, nthread=r["nthread"])
        print(key, out[key])

# --- stats kernel
cfg = benchmarks.get("cube60b", 1e6)
cfg["stats"] = 1
r = engine.run_prepared(hostcfg.prepare(cfg))
out["cube60b_stats"] = {k: v / 1e6 for k, v in r["stats"].items()}
print(out["cube60b_stats"])

# --- L2 RED microbenchmark
ms = C.c_float()
ops = C.c_uint64()
red = {}
for eb in (4, 8):
    for span in (216000, 7109137, 64 * 1024 * 1024):
        for hot in (0, 100):
            abi.check(lib.mcxb_bench_red(0, eb, span, 148 * 8, 2000, hot, 3, C.byref(ms), C.byref(ops)))
            red["%d_%d_%d" % (eb, span, hot)] = ops.value / ms.value / 1e6
            print("red", eb, span, hot, "Gops/s", ops.value / ms.value / 1e6)
out["red_gops"] = red
json.dump(out, open("gpurun_out/first.json", "w"), indent=1, default=str)
