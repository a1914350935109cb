"""Fuzz of the plain-C restatement (oracle/libmcxoracle.so) against the reference's own kernel source built for the host
(oracle/_ref): random combinations of source type, boundary codes, media, gates, detector flags and physics modes, every
output compared BIT FOR BIT.  CPU only (test infrastructure).

    python tools/oracle_fuzz.py [cases] [seed]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import decks                                              # noqa: E402
from mcxcl_b200 import benchmarks, hostcfg               # noqa: E402
from oracle import loader                                 # noqa: E402


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def random_case(rs):
    mode = rs.choice(["plain", "plain", "rf", "polarised", "traj", "continuous", "svmc", "multisrc", "adjoint"])
    name = rs.choice(sorted(decks.SOURCES))
    base = rs.choice(["cube", "two_layer"])
    if mode in ("svmc", "adjoint"):            # the reference builds of MED_TYPE 97 exist for the pencil source only
        name, base = "pencil", "cube"
    src = dict(decks.SOURCES[name])
    cfg = decks.cube(int(rs.choice([500, 1000, 1501]))) if base == "cube" else decks.two_layer(int(rs.choice([500, 1000])))
    cfg.update(src)
    tag = [base, name]
    bname = rs.choice(sorted(decks.BOUNDARIES))
    b = dict(decks.BOUNDARIES[bname])
    b.pop("nphoton", None)
    if "prop" in b and base != "cube":
        b.pop("prop")
        b["bc"] = b["bc"].replace("m", "r")
    cfg.update(b)
    tag.append(bname)
    if base == "two_layer" and not cfg.get("isreflect"):
        # label-0 voxels INSIDE the grid with the reflection code compiled out: the reference lets packets live on in them
        # (src/mcx_core.cl:2816, 2927-2929) and then indexes its per-medium rows with 0 - 1 (:2515, 2787) or divides 0 by 0
        # under gscatter -- undefined there (it crashes or never ends on the host), so the pocket is filled
        vol = cfg["vol"].copy()
        vol[vol == 0] = 3
        cfg["vol"] = vol
    if rs.rand() < 0.5:
        cfg.update(tend=2e-9, tstep=float(rs.choice([2e-10, 5e-10])))
        tag.append("gates")
    if rs.rand() < 0.6:
        flags = "".join(f for f in "dspmxvw" if rs.rand() < 0.6) or "dp"
        cfg.update(issavedet=1, savedetflag=flags, detpos=[[29, 29, 0, 6], [40, 20, 0, 3]], maxdetphoton=4000)
        tag.append("det:" + flags)
        if rs.rand() < 0.3:
            cfg.update(issaveseed=1)
            tag.append("seeds")
    else:
        cfg.update(issavedet=0)
    if rs.rand() < 0.3:
        cfg.update(outputtype=str(rs.choice(["fluence", "energy", "length"])))
        tag.append(cfg["outputtype"])
    if rs.rand() < 0.2:
        cfg.update(minenergy=0.01)
        tag.append("roulette")
    if rs.rand() < 0.15:
        cfg.update(gscatter=3)
        tag.append("gscatter")
    pattern = cfg.get("srctype") in ("pattern", "pattern3d")
    if mode == "rf" and not pattern:
        cfg.update(omega=2 * np.pi * float(rs.choice([50e6, 200e6])))
        tag.append("rf")
    elif mode == "polarised" and base == "cube" and "prop" not in b:
        nmed = len(cfg["prop"]) - 1
        cfg.update(smatrix=decks.rayleigh(nmed), srciquv=[1, float(rs.choice([0, 1])), 0, float(rs.choice([0, 1]))])
        if cfg.get("issavedet"):
            cfg["savedetflag"] = cfg["savedetflag"] + "i"
        tag.append("polarised")
    elif mode == "traj":
        cfg.update(debuglevel="M", maxjumpdebug=100000)
        tag.append("traj")
    elif mode == "continuous" and base == "cube" and "prop" not in b and not cfg.get("isspecular"):
        fmt = rs.choice(sorted(decks.media_volumes()))
        vol, prop, _ = decks.media_volumes()[fmt]
        cfg.update(vol=vol, prop=prop)
        if cfg.get("issavedet"):
            cfg["savedetflag"] = "".join(f for f in cfg["savedetflag"] if f not in "spm") or "d"
        tag.append(fmt)
    elif mode == "svmc" and base == "cube" and "prop" not in b and not cfg.get("isspecular") and name == "pencil":
        import test_gpu_svmc as sv
        cfg.update(vol=sv.tilted_slab(z0=float(rs.uniform(10, 30)), sx=float(rs.uniform(-0.3, 0.3)), sy=float(rs.uniform(-0.3, 0.3))),
                   prop=[[0, 0, 1, 1], [0.02, 1.0, 0.8, 1.37], [0.005, 2.0, 0.9, float(rs.choice([1.37, 1.5]))]], srcpos=[20, 20, 0])
        if cfg.get("issavedet"):
            cfg.update(detpos=[[20, 20, 0, 6], [30, 10, 0, 3]])
        tag.append("svmc")
    elif mode in ("multisrc", "adjoint") and not pattern and base == "cube":
        # extra sources of the same type: the main source's parameters with shifted positions
        pos = list(cfg.get("srcpos", [29.0, 29.0, 0.0]))[:3]
        cfg["srcpos"] = [pos + [1.0], [pos[0] + 6.0, pos[1] - 5.0, pos[2], 1.0], [pos[0] - 7.0, pos[1] + 4.0, pos[2], 0.5]]
        for key in ("srcdir", "srcparam1", "srcparam2"):
            if key in cfg:
                row = list(cfg[key]) + [0.0] * (4 - len(cfg[key]))
                cfg[key] = [row, row, row]
        cfg["srcid"] = int(rs.choice([0, -1, 2]))
        tag.append("multisrc srcid=%d" % cfg["srcid"])
        if mode == "adjoint" and cfg["srcid"] == -1 and name == "pencil":
            cfg.update(outputtype="adjoint", detpos=[[pos[0] - 7.0, pos[1] + 4.0, pos[2], 3.0]], srcparam1=[[0, 0, 0, 0], [0, 0, 0, 0], [3.0, 0, 0, 0]])
            tag.append("adjoint")
    return cfg, " ".join(tag)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rs = np.random.RandomState(seed)
    ref, port = loader.ref(), loader.port()
    bad = skipped = 0
    for k in range(n):
        cfg, tag = random_case(rs)
        try:
            p = hostcfg.prepare(cfg)
        except Exception as e:
            skipped += 1
            continue
        work = int(rs.choice([16, 64, 100]))
        if os.environ.get('FUZZ_VERBOSE'):
            print('case %d: %s (work %d)' % (k, tag, work), flush=True)
        try:
            a = ref.run(p, work, hostthreads=1, want_energy=True)
        except RuntimeError as e:
            skipped += 1          # the reference build for this combination does not exist (e.g. continuous media + a non-pencil source)
            continue
        b = port.run(p, work, hostthreads=1, want_energy=True)
        same = (a["energytot"] == b["energytot"] and a["energyesc"] == b["energyesc"] and (bits(a["energy"]) == bits(b["energy"])).all()
                and (bits(a["field"]) == bits(b["field"])).all() and a["detected"] == b["detected"]
                and (a["detp"] is None or (bits(a["detp"]) == bits(b["detp"])).all())
                and (a["seeds"] is None or (a["seeds"] == b["seeds"]).all())
                and (a["traj"] is None or (a["traj"].shape == b["traj"].shape and (bits(a["traj"]) == bits(b["traj"])).all()))
                and (a["n_segment"], a["n_deposit"], a["n_scatter"]) == (b["n_segment"], b["n_deposit"], b["n_scatter"]))
        if not same:
            bad += 1
            print("DIFFER case %d: %s (work %d): absorbed %.6f vs %.6f, detected %d vs %d" % (k, tag, work, a["absorbed"], b["absorbed"], a["detected"], b["detected"]))
    print("%d cases, %d skipped, %d differ" % (n, skipped, bad))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
