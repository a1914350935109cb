import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from mcxcl_b200 import benchmarks, engine, hostcfg
from oracle import loader
from util import run_gpu, run_ref
ref = loader.ref()
for me in (0.0, 0.01):
    cfg = benchmarks.get("skinvessel", 100000)
    cfg["minenergy"] = me
    p, r = run_gpu(cfg)
    _, o = run_ref(ref, cfg)
    lab = (p.keep["vol"] & 0x7FFFFFFF).ravel()
    g = r["field"].astype(np.float64) / r["normalizer"]
    w = o["field"].astype(np.float64)
    print("minenergy", me, "absorbed gpu/ref", r["absorbed"], o["absorbed"])
    for m in range(1, 5):
        mua = float(p.keep["prop"][m, 0])
        print("  label", m, "mua", mua, "sum gpu", g[lab == m].sum(), "ref", w[lab == m].sum(), "ratio", g[lab == m].sum() / w[lab == m].sum(),
              "energy gpu", g[lab == m].sum() * mua, "ref", w[lab == m].sum() * mua)
