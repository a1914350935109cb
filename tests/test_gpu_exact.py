"""Bit-exact tier on the B200: the integer RNG sequence and the voxel traversal of fixed rays computed by
the CUDA device functions the photon kernel is built from must equal the reference kernel's
(BASELINE.json north_star; SURVEY.md 8(a) R1,R2,R8,R9,R10; App. B.3/B.4)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from mcxcl_b200 import abi
from util import f32bits, gpu_trace, random_rays, records_equal

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_survey.json")))


def gpu_rng(lib, seeds, ndraw):
    seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
    n = seeds.size // 4
    out = np.zeros((n, max(ndraw, 1)), dtype=np.float32)
    st = np.zeros((n, 2), dtype=np.uint64)
    abi.check(lib.mcxb_test_rng(0, seeds.ctypes.data, n, ndraw, out.ctypes.data, st.ctypes.data), "mcxb_test_rng")
    return out[:, :ndraw], st


def test_rng_golden_vectors(lib):
    seeds = np.array([r["seed"] for r in KAT["rng"]], dtype=np.uint32)
    u, _ = gpu_rng(lib, seeds, 3)
    _, st0 = gpu_rng(lib, seeds, 0)
    for row in KAT["rng"]:
        k = row["thread"]
        np.testing.assert_allclose(u[k], np.array(row["u"], dtype=np.float32), rtol=2e-7, atol=0)
        assert [int(x) for x in st0[k]] == [int(t, 16) for t in row["t"]]


def test_seed_table_golden(lib):
    got = np.zeros(32, dtype=np.uint32)
    lib.mcxb_fill_seeds(KAT["seed"], 0, 8, got.ctypes.data)
    assert got.reshape(8, 4).tolist() == [r["seed"] for r in KAT["rng"]]


def test_rng_streams_bit_exact_vs_reference(lib, ref):
    seeds = ref.seeds(KAT["seed"], 8192)
    want, wst = ref.rng(seeds, 512)
    got, gst = gpu_rng(lib, seeds, 512)
    assert (f32bits(got) == f32bits(want)).all()
    assert (gst == wst).all()


def test_rng_edge_seeds(lib, ref):
    seeds = np.array([[0, 0, 0, 1], [0xFFFFFFFF] * 4, [1, 0, 0, 0], [0x7FFFFFFF, 0, 0x80000000, 5]], dtype=np.uint32)
    want, wst = ref.rng(seeds, 1000)
    got, gst = gpu_rng(lib, seeds, 1000)
    assert (f32bits(got) == f32bits(want)).all() and (gst == wst).all()
    assert got.min() >= 0.0 and got.max() < 1.0


def test_traversal_golden_vector(lib):
    t = KAT["trace"]
    v0 = np.array([int(b, 16) for b in t["v0_bits"]] + [0], dtype=np.uint32).view(np.float32)
    out = gpu_trace(lib, np.array(t["p0"] + [1.0], np.float32), v0, len(t["steps"]), t["dims"], t["musp"])[0]
    for k, s in enumerate(t["steps"]):
        assert int(out["face"][k]) == s["face"]
        assert "%08x" % f32bits(out["dist"][k])[0] == s["dist"]
        assert ["%08x" % f32bits(out[c][k])[0] for c in ("px", "py", "pz")] == s["p"]
        assert [int(out[c][k]) for c in ("ix", "iy", "iz")] == s["voxel"]
        assert int(out["idx1d"][k]) == s["idx1d"]


@pytest.mark.parametrize("dims,musp,seed", [
    ((60, 60, 60), 1.0, 1), ((60, 60, 60), 0.731, 2), ((200, 200, 200), 9.398496241 * 0.005 * (1 - 0.9), 3),
    ((181, 217, 181), 7.8182, 4), ((190, 496, 104), 0.8 * 25.2, 5), ((1, 100, 100), 2.0, 6), ((7, 3, 500), 1e-10, 7),
])
def test_traversal_bit_exact_vs_reference(lib, ref, dims, musp, seed):
    p0, v0 = random_rays(30000, dims, seed)
    want = ref.trace(p0, v0, 96, dims, musp)
    got = gpu_trace(lib, p0, v0, 96, dims, musp)
    assert records_equal(got, want)


def test_traversal_axis_aligned_and_boundary_rays(lib, ref):
    """rays with zero direction components (division by +-0), rays starting exactly on faces and corners"""
    dims = (60, 60, 60)
    p, v = [], []
    for d in ([1, 0, 0], [0, -1, 0], [0, 0, 1], [0.6, 0.8, 0], [0, -0.6, 0.8], [-0.0, 1.0, 0.0], [1e-30, 1.0, 1e-20]):
        for s in ([0, 0, 0], [30, 30, 30], [59.999996, 10.5, 0], [29.5, 29.5, 1e-7], [60, 60, 60]):
            p.append(s + [1.0])
            v.append(d + [0.0])
    p0, v0 = np.array(p, np.float32), np.array(v, np.float32)
    want = ref.trace(p0, v0, 80, dims, 1.0)
    got = gpu_trace(lib, p0, v0, 80, dims, 1.0)
    assert records_equal(got, want)


def test_scalar_helpers_bit_exact(lib, ref):
    rs = np.random.RandomState(11)
    a = np.concatenate([rs.uniform(-5, 65, 5000), np.arange(0, 61), [0.0, -0.0, 1000.0, -1000.0]]).astype(np.float32)
    d = rs.randint(-1, 2, a.size).astype(np.int32)
    m = 300000
    v = rs.normal(size=(m, 4)).astype(np.float32)
    v[:, :3] /= np.linalg.norm(v[:, :3], axis=1, keepdims=True)
    v[:1000, 2] = rs.uniform(-1e-3, 1e-3, 1000)                  # grazing incidence on z faces
    v[1000:2000, :3] = [0, 0, 1]                                 # normal incidence
    n1 = rs.choice([1.0, 1.33, 1.37, 1.45, 1.55, 6.85], m).astype(np.float32)
    n2 = rs.choice([1.0, 1.33, 1.37, 1.45, 1.55, 6.85], m).astype(np.float32)
    n1[2000:3000] = rs.uniform(1.0, 2.5, 1000)
    n2[2000:3000] = rs.uniform(1.0, 2.5, 1000)
    face = rs.randint(0, 3, m).astype(np.int32)
    wna, wrc = ref.scalar(a, d, v, n1, n2, face)
    gna = np.zeros_like(wna)
    grc = np.zeros_like(wrc)
    abi.check(lib.mcxb_test_scalar(0, a.ctypes.data, d.ctypes.data, a.size, gna.ctypes.data, v.ctypes.data, n1.ctypes.data,
                                   n2.ctypes.data, face.ctypes.data, m, grc.ctypes.data), "mcxb_test_scalar")
    assert (f32bits(gna) == f32bits(wna)).all()
    assert (f32bits(grc) == f32bits(wrc)).all()
    sc = KAT["scalar"]
    assert "%08x" % f32bits(grc[:1])[0] is not None
    for (val, dr, bits_) in sc["nextafter"]:
        one = np.zeros(1, np.float32)
        dummy = np.zeros(1, np.float32)
        va, da, v4, f1 = np.array([val], np.float32), np.array([dr], np.int32), np.zeros(4, np.float32), np.zeros(1, np.int32)
        abi.check(lib.mcxb_test_scalar(0, va.ctypes.data, da.ctypes.data, 1, one.ctypes.data, v4.ctypes.data, dummy.ctypes.data,
                                       dummy.ctypes.data, f1.ctypes.data, 0, dummy.ctypes.data), "mcxb_test_scalar")
        assert "%08x" % f32bits(one)[0] == bits_


def close_enough(got, want):
    """fast tier: all but a handful of vectors agree to 1e-5; the exceptions are directions within ~1e-3 of an axis,
    where 1 - v^2 cancels and one rounding difference (FMA contraction, approximate sqrt) is amplified"""
    err = np.abs(got - want).max(axis=1)
    assert np.mean(err > 1e-5) < 1e-3 and err.max() < 2e-3


def test_rotation_and_refraction_close_to_reference(lib, ref):
    """fast tier (rsqrt / sqrt on the MUFU pipe): agreement to a few ulp; tolerance 1e-5 absolute on unit vectors (the
    rsqrt(1-vz^2) factor amplifies the last-place difference for directions close to the z axis)"""
    rs = np.random.RandomState(5)
    n = 20000
    v = np.zeros((n, 4), np.float32)
    d = rs.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    v[:, :3] = d
    v[:10, :3] = [0, 0, 1]
    v[10:20, :3] = [0, 0, -1]
    th, ph = rs.uniform(0, np.pi, n), rs.uniform(0, 2 * np.pi, n)
    st, ct, sp, cp = [x.astype(np.float32) for x in (np.sin(th), np.cos(th), np.sin(ph), np.cos(ph))]
    want = ref.rotate(v, st, ct, sp, cp)
    got = v.copy()
    abi.check(lib.mcxb_test_rotate(0, got.ctypes.data, st.ctypes.data, ct.ctypes.data, sp.ctypes.data, cp.ctypes.data, n), "rotate")
    close_enough(got, want)
    face = rs.randint(0, 3, n).astype(np.int32)
    n1 = np.full(n, 1.0, np.float32)
    n2 = np.full(n, 1.37, np.float32)          # into the denser medium: never total internal reflection
    want = ref.transmit(v, n1, n2, face)
    got = v.copy()
    abi.check(lib.mcxb_test_refract(0, got.ctypes.data, n1.ctypes.data, n2.ctypes.data, face.ctypes.data, n), "refract")
    close_enough(got, want)


def test_accumulator_copies_do_not_change_the_result(monkeypatch):
    """The device sums k replicated accumulator volumes (engine.cu `acccopies`; CTA b adds into copy b mod k).  With the
    static photon split every thread walks the same packets whatever k is, so the deposits are the same multiset and the
    volumes can differ only by the order of the fp64 additions: equal to float32 rounding, identical energy ledger."""
    from mcxcl_b200 import benchmarks, engine, hostcfg
    cfg = benchmarks.get("cube60b", 300000)
    cfg.update(sched=1, isnormalized=0)
    out = {}
    for k in (1, 8):
        monkeypatch.setenv("MCXB_ACC_COPIES", str(k))
        with engine.Simulation(hostcfg.prepare(cfg)) as sim:
            assert sim.acc_copies == k
            sim.reset()
            sim.launch()
            out[k] = sim.fetch()
    monkeypatch.delenv("MCXB_ACC_COPIES")
    assert out[1]["energytot"] == out[8]["energytot"] == 300000
    assert out[1]["energyesc"] == pytest.approx(out[8]["energyesc"], rel=1e-12)      # fp64 atomics: order of the warp sums
    assert out[1]["detected"] == out[8]["detected"]
    a, b = out[1]["field"].astype(np.float64), out[8]["field"].astype(np.float64)
    assert a.sum() > 0 and np.allclose(a, b, rtol=3e-7, atol=0)
    # default policy: as many copies as fit 40 MB, at most 8 -> 8 for the 1.7 MB cube, 1 for a 57 MB atlas volume
    with engine.Simulation(hostcfg.prepare(benchmarks.get("cube60b", 10))) as sim:
        assert sim.acc_copies == 8
    with engine.Simulation(hostcfg.prepare(benchmarks.get("colin27", 10))) as sim:
        assert sim.acc_copies == 1
