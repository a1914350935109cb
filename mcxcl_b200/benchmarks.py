"""Built-in simulation decks of the reference that BASELINE.json quotes the metric on.

Each function returns a dict in ``pmcxcl.run(**cfg)`` form.  Definitions follow the reference's
embedded benchmark JSON (src/mcx_bench.h) and example decks:
  cube60 / cube60b      src/mcx_bench.h:33-140  (example/benchmark/benchmark1.json, -b 0 / -b 1)
  cube60planar          src/mcx_bench.h:143-200
  qtest                 example/quicktest/qtest.inp:1-16
  skinvessel            src/mcx_bench.h:364-428 (example/skinvessel/mcxyz_bench.json)
  colin27               src/mcx_bench.h:572-626 (volume: tests/golden/volumes, see make_volumes.py)
  digimouse             example/digimouse/digimouse.json (fourier source as shipped) and the
                        multi-source, time-gated variant BASELINE.json names
"""
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
VOLUME_DIR = os.path.join(os.path.dirname(_HERE), "tests", "golden", "volumes")
BENCH_SEED = 1648335518

_CUBE_DETS = [[29.0, 19.0, 0.0, 1.0], [29.0, 39.0, 0.0, 1.0], [19.0, 29.0, 0.0, 1.0], [39.0, 29.0, 0.0, 1.0]]
_CUBE_PROP = [[0, 0, 1, 1], [0.005, 1.0, 0.01, 1.37], [0.002, 5.0, 0.9, 1.0]]


def _cube_base(nphoton):
    return dict(nphoton=int(nphoton), vol=np.ones((60, 60, 60), dtype=np.uint8), prop=_CUBE_PROP,
                tstart=0.0, tend=5e-9, tstep=5e-9, seed=BENCH_SEED, issrcfrom0=1,
                srcpos=[29.0, 29.0, 0.0], srcdir=[0.0, 0.0, 1.0], detpos=_CUBE_DETS)


def cube60(nphoton=1e6):
    cfg = _cube_base(nphoton)
    cfg.update(session="cube60", isreflect=0)
    return cfg


def cube60b(nphoton=1e6):
    cfg = _cube_base(nphoton)
    cfg.update(session="cube60b", isreflect=1)
    return cfg


def cube60planar(nphoton=1e6):
    cfg = _cube_base(nphoton)
    cfg.update(session="cube60planar", isreflect=1, srctype="planar", srcpos=[10.0, 10.0, -10.0],
               srcparam1=[40.0, 0.0, 0.0, 0.0], srcparam2=[0.0, 40.0, 0.0, 0.0])
    return cfg


def qtest(nphoton=1e6):
    """example/quicktest/qtest.inp: medium 1 = mus 1, g 0.01, mua 0.005, n 1 (order mus g mua n)."""
    cfg = _cube_base(nphoton)
    cfg.update(session="qtest", isreflect=1, issrcfrom0=0, srcpos=[30.0, 30.0, 1.0],
               prop=[[0, 0, 1, 1], [0.005, 1.0, 0.01, 1.0]],
               detpos=[[30.0, 20.0, 1.0, 1.0], [30.0, 40.0, 1.0, 1.0], [20.0, 30.0, 1.0, 1.0], [40.0, 30.0, 1.0, 1.0]])
    return cfg


def skinvessel_volume():
    """Grid(tag 1) + ZLayers [[1,20,1],[21,32,4],[33,200,3]] + Cylinder(tag 2, R 20) along x.

    Rasterised like the reference's shape parser with OriginType 1: a voxel (i,j,k) belongs to the
    cylinder when its centre (i+.5, j+.5, k+.5) lies within R of the axis (src/mcx_shapes.c)."""
    vol = np.ones((200, 200, 200), dtype=np.uint8)
    vol[:, :, 0:20] = 1
    vol[:, :, 20:32] = 4
    vol[:, :, 32:200] = 3
    c = np.arange(200, dtype=np.float64) + 0.5
    yy, zz = np.meshgrid(c, c, indexing="ij")
    mask = (yy - 100.5) ** 2 + (zz - 100.5) ** 2 <= 20.0 ** 2
    vol[:, mask] = 2
    return vol


def skinvessel(nphoton=1e6):
    return dict(session="skinvessel", nphoton=int(nphoton), vol=skinvessel_volume(), unitinmm=0.005,
                prop=[[1e-05, 0, 1, 1.37], [3.564e-05, 1, 1, 1.37], [23.05426549, 9.398496241, 0.9, 1.37],
                      [0.04584957865, 35.65405549, 0.9, 1.37], [1.657237447, 37.59398496, 0.9, 1.37]],
                tstart=0.0, tend=5e-8, tstep=5e-8, seed=BENCH_SEED, issrcfrom0=1, isreflect=1,
                srctype="disk", srcpos=[100.0, 100.0, 20.0], srcdir=[0.0, 0.0, 1.0], srcparam1=[60.0, 0.0, 0.0, 0.0],
                issavedet=0)


def load_volume(name):
    path = os.path.join(VOLUME_DIR, name + ".npz")
    if not os.path.exists(path):
        raise FileNotFoundError("%s missing: regenerate with tests/golden/make_volumes.py" % path)
    return np.load(path)["vol"]


def colin27(nphoton=1e6):
    return dict(session="colin27", nphoton=int(nphoton), vol=load_volume("colin27"),
                prop=[[0.0, 0.0, 1.0, 1.0], [0.019, 7.8182, 0.89, 1.37], [0.019, 7.8182, 0.89, 1.37],
                      [0.0004, 0.009, 0.89, 1.37], [0.02, 9.0, 0.89, 1.37], [0.08, 40.9, 0.84, 1.37],
                      [0.0, 0.0, 1.0, 1.0]],
                tstart=0.0, tend=5e-9, tstep=5e-9, seed=BENCH_SEED, issrcfrom0=1, isreflect=1,
                srcpos=[75.0, 67.38, 167.5], srcdir=[0.1636, 0.4569, -0.8743],
                detpos=[[75.0, 77.19, 170.3, 1.0], [75.0, 89.0, 171.6, 1.0], [75.0, 97.67, 172.4, 1.0], [75.0, 102.4, 172.0, 1.0]])


_DIGIMOUSE_PROP = None


def digimouse(nphoton=1e6, variant="shipped"):
    """variant 'shipped': the single fourier widefield source of example/digimouse/digimouse.json.
    variant 'multisrc_tg': the BASELINE.json reading -- 4 pencil sources (srcid=-1, one output volume
    per source) and 10 time gates of 0.5 ns; defined in DESIGN.md section 7."""
    d = np.load(os.path.join(VOLUME_DIR, "digimouse.npz"))
    cfg = dict(session="digimouse", nphoton=int(nphoton), vol=d["vol"], prop=d["prop"], unitinmm=float(d["unitinmm"]),
               tstart=0.0, tend=5e-9, tstep=5e-9, seed=BENCH_SEED, issrcfrom0=1, isreflect=1, issavedet=0)
    if variant == "shipped":
        cfg.update(srctype="fourier", srcpos=[50.0, 200.0, 100.0], srcdir=[0.0, 0.0, -1.0],
                   srcparam1=[100.0, 0.0, 0.0, 2.0], srcparam2=[0.0, 100.0, 0.0, 0.0])
    elif variant == "multisrc_tg":
        cfg.update(srctype="pencil", srcid=-1, tstep=5e-10,
                   srcpos=[[95.0, 150.0, 100.0, 1.0], [95.0, 250.0, 100.0, 1.0], [95.0, 350.0, 100.0, 1.0], [95.0, 420.0, 100.0, 1.0]],
                   srcdir=[[0.0, 0.0, -1.0, 0.0]] * 4)
    else:
        raise ValueError(variant)
    return cfg


def digimouse_tg(nphoton=1e6):
    """BASELINE.json's "multi-source time-gated" digimouse: 4 sources x 10 gates x 9.8 M voxels"""
    return digimouse(nphoton, "multisrc_tg")


BENCHMARKS = {"cube60": cube60, "cube60b": cube60b, "cube60planar": cube60planar, "qtest": qtest,
              "skinvessel": skinvessel, "colin27": colin27, "digimouse": digimouse, "digimouse_tg": digimouse_tg}


def get(name, nphoton=1e6, **kw):
    return BENCHMARKS[name](nphoton, **kw)
