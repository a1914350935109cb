"""Extended physics: polarised light, RF (frequency-domain) runs, adjoint products.

Reference: Stokes vectors and Mie phase functions src/mcx_core.cl:792-835, 1633-1638, 2454-2468, 2564; complex packet
weights of an RF forward run :2427-2430, 2750-2763, 2833-2841, 3035-3040; RF replay Jacobians :2257-2263, 2572-2577,
2853-2855, 2895-2897; host side src/mcx_host.cpp:473, 770, 1243-1276, 1389-1396; adjoint post-kernels
src/mcx_core.cl:3313-3512 launched by src/mcx_host.cpp:1468-1641.

Nothing in the reference's own tests pins these modes; the checker is the reference's kernel source built for the host
(oracle/_ref) run on the same decks, plus identities that hold by construction."""
import ctypes as C

import numpy as np
import pytest

import decks
from mcxcl_b200 import abi, benchmarks, engine, hostcfg
from test_replay import replay_cfg

OMEGA = 2 * np.pi * 100e6       # 100 MHz modulation


rayleigh, isotropic_matrix, pol_cfg = decks.rayleigh, decks.isotropic_matrix, decks.pol_cfg


# ------------------------------------------------------------------------------------------- CPU: host logic and the oracle
def test_host_fields_of_the_extended_modes():
    p = hostcfg.prepare(pol_cfg(10))
    assert p.c.polmedianum == 1 and p.c.savedetflag & 0x80 and p.reclen == 1 + 1 + 3 + 3 + 1 + 4
    assert hostcfg.prepare(dict(pol_cfg(10), smatrix=None)).c.savedetflag & 0x80 == 0         # src/mcx_utils.c:1777-1781
    with pytest.raises(hostcfg.ConfigError):
        hostcfg.prepare(pol_cfg(10, smatrix=rayleigh(2)))                                   # one matrix per medium (:1540-1542)
    p = hostcfg.prepare(dict(benchmarks.get("cube60", 10), omega=OMEGA))
    assert p.rfplanes == 2 and p.fieldlen == 2 * 216000                                     # src/pmcxcl.cpp:1209-1211
    assert hostcfg.prepare(dict(benchmarks.get("cube60", 10), outputtype="adjoint")).c.outputtype == 11
    with pytest.raises(hostcfg.ConfigError):
        hostcfg.prepare(dict(benchmarks.get("cube60", 10), outputtype="rf"))                # an RF Jacobian is a replay output


def test_oracle_polarised_run_is_sane(ref):
    """the reference source with a Mueller matrix: every detected Stokes vector is normalised to I = 1 with |Q,U,V| <= 1, and
    with the matrix of an isotropic, non-polarising scatterer the run is the g = 0 run of the unpolarised code"""
    o = ref.run(hostcfg.prepare(pol_cfg(20000)), 256, hostthreads=0)
    s = o["detp"][:, -4:]
    assert o["detected"] > 50 and np.all(s[:, 0] == 1.0) and np.all(np.abs(s[:, 1:]) <= 1.0 + 1e-5)
    assert np.all(s[:, 1] ** 2 + s[:, 2] ** 2 + s[:, 3] ** 2 <= 1.0 + 1e-4)
    iso = ref.run(hostcfg.prepare(pol_cfg(40000, smatrix=isotropic_matrix(1), issavedet=0)), 256, hostthreads=0)
    plain = ref.run(hostcfg.prepare(pol_cfg(40000, smatrix=None, issavedet=0)), 256, hostthreads=0)
    assert abs(iso["absorbed"] - plain["absorbed"]) < 0.01


def test_oracle_rf_forward_limits(ref):
    """RF forward in the reference source: with a vanishing modulation frequency the real part is the ordinary fluence and
    the imaginary part vanishes; at 100 MHz the phase lag makes the imaginary part negative and |field| smaller"""
    base = dict(benchmarks.get("cube60b", 20000), issavedet=0, isnormalized=0)
    plain = ref.run(hostcfg.prepare(base), 256, hostthreads=0)["field"].astype(np.float64)
    slow = ref.run(hostcfg.prepare(dict(base, omega=1.0)), 256, hostthreads=0)["field"].astype(np.float64)
    assert slow.size == 2 * plain.size
    assert slow[:plain.size].sum() == pytest.approx(plain.sum(), rel=1e-4) and abs(slow[plain.size:].sum()) < 1e-6 * plain.sum()
    fast = ref.run(hostcfg.prepare(dict(base, omega=OMEGA)), 256, hostthreads=0)["field"].astype(np.float64)
    re, im = fast[:plain.size].sum(), fast[plain.size:].sum()
    assert 0 < re < plain.sum() and im < 0 and np.hypot(re, im) < plain.sum()


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_polarised_run_matches_the_reference_source(ref):
    cfg = pol_cfg(1000000)
    p = hostcfg.prepare(cfg)
    with engine.Simulation(p) as sim:
        assert sim.kernel_name.endswith("/ext")          # the extended-physics specialisation was selected
    g = engine.run_prepared(p)
    o = ref.run(hostcfg.prepare(pol_cfg(300000)), 4096, hostthreads=0)
    assert g["energytot"] == 1000000
    assert abs(g["absorbed"] - o["absorbed"]) < 4 * 1.6 * np.sqrt(0.27 * 0.73 / 300000)
    rate_g, rate_o = g["detected"] / 1e6, o["detected"] / 3e5
    assert abs(rate_g - rate_o) < 5 * np.sqrt(rate_o / 3e5)
    sg, so = g["detp"][:, -4:].astype(np.float64), o["detp"][:, -4:].astype(np.float64)
    assert np.all(sg[:, 0] == 1.0) and np.all(np.abs(sg[:, 1:]) <= 1.0 + 1e-4)
    # degree of polarisation left in the detected light: means and spreads of Q, U, V agree
    for k in (1, 2, 3):
        se = np.sqrt(so[:, k].var() / len(so) + sg[:, k].var() / len(sg)) + 1e-9
        assert abs(sg[:, k].mean() - so[:, k].mean()) < 5 * se, k
        assert sg[:, k].std() == pytest.approx(so[:, k].std(), rel=0.15, abs=0.01), k
    assert np.abs(sg[:, 1:3]).mean() > 0.2          # the detected light IS partially polarised (otherwise the test says nothing)
    # fluence
    fg, fo = g["field"].astype(np.float64), o["field"].astype(np.float64) / (o["energytot"] * 5e-9)      # flux normalisation, src/mcx_host.cpp:1389-1391
    assert fg.sum() == pytest.approx(fo.sum(), rel=0.01)
    # the isotropic, non-polarising matrix reproduces the unpolarised g = 0 run
    iso = engine.run_prepared(hostcfg.prepare(pol_cfg(1000000, smatrix=isotropic_matrix(1), issavedet=0)))
    plain = engine.run_prepared(hostcfg.prepare(pol_cfg(1000000, smatrix=None, issavedet=0)))
    assert abs(iso["absorbed"] - plain["absorbed"]) < 0.006          # two 1e6-photon runs: sigma of the difference 1e-3


@pytest.mark.gpu
def test_gpu_rf_forward_matches_the_reference_source(ref):
    base = dict(benchmarks.get("cube60b", 1000000), issavedet=0, omega=OMEGA)
    p = hostcfg.prepare(base)
    g = engine.run_prepared(p)
    o = ref.run(hostcfg.prepare(dict(base, nphoton=300000)), 4096, hostthreads=0)
    n = 216000
    assert g["field"].size == 2 * n
    gre, gim = g["field"][:n].astype(np.float64), g["field"][n:].astype(np.float64)
    ore, oim = o["field"][:n].astype(np.float64), o["field"][n:].astype(np.float64)
    # the checker returns raw deposits; normalise them the way the engine did (src/mcx_host.cpp:1389-1396, flux: 1/(E tstep))
    scale = 1.0 / (o["energytot"] * 5e-9)
    ore, oim = ore * scale, oim * scale
    assert abs(g["absorbed"] - o["absorbed"]) < 4 * 1.6 * np.sqrt(0.27 * 0.73 / 300000)
    assert gre.sum() == pytest.approx(ore.sum(), rel=0.015) and gim.sum() == pytest.approx(oim.sum(), rel=0.04)
    assert gim.sum() < 0 < gre.sum()
    # phase and amplitude along the beam axis, away from the noise floor
    # (a 5 x 5 column around the axis, so that one run's voxel noise does not decide the test)
    vol_g = (gre + 1j * gim).reshape(60, 60, 60)[:, 27:32, 27:32].sum(axis=(1, 2))
    vol_o = (ore + 1j * oim).reshape(60, 60, 60)[:, 27:32, 27:32].sum(axis=(1, 2))
    sel = slice(1, 22)
    np.testing.assert_allclose(np.abs(vol_g[sel]), np.abs(vol_o[sel]), rtol=0.1)
    assert np.max(np.abs(np.angle(vol_g[sel] / vol_o[sel]))) < 0.05
    flux = engine.run(base)["flux"]
    assert flux.dtype == np.complex64 and flux.shape == (60, 60, 60, 1)
    # a vanishing frequency gives back the ordinary run
    slow = engine.run_prepared(hostcfg.prepare(dict(base, omega=1.0)))
    plain = engine.run_prepared(hostcfg.prepare(dict(base, omega=0.0)))
    assert slow["field"][:n].astype(np.float64).sum() == pytest.approx(plain["field"].astype(np.float64).sum(), rel=5e-3)
    assert abs(slow["field"][n:].astype(np.float64).sum()) < 1e-5 * plain["field"].astype(np.float64).sum()


def test_reference_rf_replay_loses_its_phase_factors(ref):
    """Why the RF replay outputs are NOT compared with the reference source: launchnewphoton advances its ppath pointer
    by partialdata (src/mcx_core.cl:1620) and then stores cos / sin(omega tof) at ppath[w0offset + srcnum] (:2261-2262),
    i.e. partialdata floats further than where the main loop reads them (:2573-2574, 2854, 2896) -- past this work-item's
    shared-memory row.  A replay always carries partial paths (mcx_replayinit insists on the P flag), so the factors are
    read as zero and the reference's RF Jacobian is empty; the absorption Jacobian of the same records is not."""
    cfg = dict(benchmarks.get("cube60", 50000), issaveseed=1, savedetflag="DSPM")
    base = ref.run(hostcfg.prepare(cfg), 1024, hostthreads=0)
    assert base["detected"] > 50
    jac = ref.run(hostcfg.prepare(replay_cfg(cfg, base["detp"], base["seeds"], outputtype="jacobian", isnormalized=0)), 1, hostthreads=0)
    rf = ref.run(hostcfg.prepare(replay_cfg(cfg, base["detp"], base["seeds"], outputtype="rf", isnormalized=0, omega=OMEGA)), 1, hostthreads=0)
    assert jac["field"].sum() > 0 and not rf["field"].any()


@pytest.mark.gpu
@pytest.mark.parametrize("ot", ["rf", "rfmus"])
def test_gpu_rf_replay_satisfies_the_identities(ot):
    """RF Jacobians of a replay (src/mcx_core.cl:2257-2263, 2572-2577, 2853-2855, 2895-2897), photon by photon:
        rf:     sum(Re) = -sum_i w_i L_i cos(omega t_i),   sum(Im) = -sum_i w_i L_i sin(omega t_i)
        rfmus:  sum(Re) =  sum_i w_i nscat_i cos(omega t_i), sum(Im) = sum_i w_i nscat_i sin(omega t_i)
    and voxel by voxel the rf output is the absorption Jacobian of the same records with every photon's deposits scaled by
    -cos / -sin(omega t_i) (checked with a replay of ONE photon at a time for a few photons)"""
    cfg = dict(benchmarks.get("cube60", 300000), issaveseed=1, savedetflag="DSPM")
    base = engine.run_prepared(hostcfg.prepare(cfg))
    n = min(1500, base["detected"])
    rc = replay_cfg(cfg, base["detp"][:n], base["seeds"][:n], outputtype=ot, isnormalized=0, omega=OMEGA)
    p = hostcfg.prepare(rc)
    assert p.rfplanes == 2
    g = engine.run_prepared(p)["field"].astype(np.float64)
    w, tof = p.keep["replay_weight"].astype(np.float64), p.keep["replay_tof"].astype(np.float64)
    d = base["detp"][:n][p.keep["replay_index"]]
    M = p.c.medianum - 1                   # record: detector id, M scattering counts (uint32 bits), M partial paths, M momentum transfers
    nscat = np.ascontiguousarray(d[:, 1:1 + M]).view(np.uint32).astype(np.float64).sum(axis=1)
    plen = d[:, 1 + M:1 + 2 * M].astype(np.float64).sum(axis=1)
    F = 216000
    if ot == "rf":
        assert g[:F].sum() == pytest.approx(-(w * plen * np.cos(OMEGA * tof)).sum(), rel=1e-3)
        assert g[F:].sum() == pytest.approx(-(w * plen * np.sin(OMEGA * tof)).sum(), rel=1e-3)
    else:
        assert g[:F].sum() == pytest.approx((w * nscat * np.cos(OMEGA * tof)).sum(), rel=1e-3)
        assert g[F:].sum() == pytest.approx((w * nscat * np.sin(OMEGA * tof)).sum(), rel=1e-3)
    base_ot = "jacobian" if ot == "rf" else "wp"
    sign = -1.0 if ot == "rf" else 1.0
    for i in (0, 7, 19):
        one = dict(rc, seed=rc["seed"][:, i:i + 1], detphotons=rc["detphotons"][:, i:i + 1])
        pj = hostcfg.prepare(dict(one, outputtype=base_ot))
        if pj.c.nphoton == 0:
            continue
        j = engine.run_prepared(pj)["field"].astype(np.float64)
        r = engine.run_prepared(hostcfg.prepare(one))["field"].astype(np.float64)
        t = float(pj.keep["replay_tof"][0])
        np.testing.assert_allclose(r[:F], sign * j * np.cos(OMEGA * t), rtol=2e-3, atol=1e-7 * np.abs(j).max())
        np.testing.assert_allclose(r[F:], sign * j * np.sin(OMEGA * t), rtol=2e-3, atol=1e-7 * np.abs(j).max())


def numpy_adjoint(re, im, dims, maxgate, ns, nd, gradient):
    """independent statement of mcx_adjoint_kernel / mcx_adjoint_dcoeff_kernel (src/mcx_core.cl:3393-3512)"""
    nx, ny, nz = dims
    n = nx * ny * nz

    def cw(f):
        return f.reshape(ns + nd, maxgate, nz, ny, nx).astype(np.float32).sum(axis=1, dtype=np.float32)

    fr = cw(re)
    fi = cw(im) if im is not None else np.zeros_like(fr)
    if gradient:
        def grad(v):
            return np.stack([np.gradient(v, axis=a, edge_order=2) if v.shape[a] > 2 else (np.diff(v, axis=a).repeat(2, axis=a) if v.shape[a] == 2 else np.zeros_like(v))
                             for a in (2, 1, 0)])       # d/dx, d/dy, d/dz
        gr, gi = [grad(v) for v in fr], [grad(v) for v in fi]
    out = np.zeros((2 if im is not None else 1, ns * nd, n), np.float64)
    for s in range(ns):
        for d in range(nd):
            if gradient:
                a, b, c, e = gr[s], gi[s], gr[ns + d], gi[ns + d]
                out[0, s * nd + d] = ((a * c).sum(0) - (b * e).sum(0)).ravel()
                if im is not None:
                    out[1, s * nd + d] = ((a * e).sum(0) + (b * c).sum(0)).ravel()
            else:
                out[0, s * nd + d] = (fr[s] * fr[ns + d] - fi[s] * fi[ns + d]).ravel()
                if im is not None:
                    out[1, s * nd + d] = (fr[s] * fi[ns + d] + fi[s] * fr[ns + d]).ravel()
    return out.ravel()


@pytest.mark.gpu
@pytest.mark.parametrize("gradient", [0, 1])
@pytest.mark.parametrize("rf", [False, True])
def test_gpu_adjoint_products(gradient, rf):
    lib = abi.load()
    rs = np.random.RandomState(7 + gradient)
    dims, maxgate, ns, nd = (9, 7, 5), 3, 2, 3
    n = dims[0] * dims[1] * dims[2]
    re = rs.uniform(0, 1, n * maxgate * (ns + nd)).astype(np.float32)
    im = rs.uniform(-1, 1, re.size).astype(np.float32) if rf else None
    out = np.zeros(n * ns * nd * (2 if rf else 1), np.float32)
    abi.check(lib.mcxb_adjoint_products(0, re.ctypes.data, im.ctypes.data if rf else None, dims[0], dims[1], dims[2], maxgate, ns, nd, gradient,
                                        out.ctypes.data), "mcxb_adjoint_products")
    want = numpy_adjoint(re, im, dims, maxgate, ns, nd, gradient)
    np.testing.assert_allclose(out, want, rtol=2e-5, atol=2e-5)


@pytest.mark.gpu
def test_gpu_adjoint_run_is_a_fluence_run_over_sources_and_detectors():
    """an adjoint output type runs the forward kernel as a fluence run with one volume per source (src/mcx_core.cl:2844,
    src/mcx_host.cpp:1389-1394); the products of the source and detector volumes are the Jacobian"""
    cfg = dict(benchmarks.get("cube60", 200000), issavedet=0, srcpos=[[30, 30, 1, 1], [30, 40, 1, 1]], srcdir=[[0, 0, 1, 0], [0, 0, 1, 0]],
               srcid=-1, tstep=2.5e-9, sched=1)         # static photon split: the same packets walk in both runs (no detectors: both in the common kernels)
    a = engine.run_prepared(hostcfg.prepare(dict(cfg, outputtype="adjoint")))
    f = engine.run_prepared(hostcfg.prepare(dict(cfg, outputtype="fluence")))
    assert a["field"].size == 216000 * 2 * 2 and a["field"].sum() > 0
    np.testing.assert_allclose(a["field"], f["field"], rtol=1e-5, atol=1e-6 * float(f["field"].max()))
    out = np.zeros(216000, np.float32)
    abi.check(abi.load().mcxb_adjoint_products(0, a["field"].ctypes.data, None, 60, 60, 60, 2, 1, 1, 0, out.ctypes.data))
    v = a["field"].astype(np.float64).reshape(2, 2, 216000).sum(axis=1)
    np.testing.assert_allclose(out, v[0] * v[1], rtol=1e-5, atol=1e-12)


@pytest.mark.gpu
def test_gpu_adjoint_launches_detectors_as_disks_like_the_reference(ref):
    """adjoint forward run (src/mcx_core.cl:2154-2183): the sources appended for the detectors start from a disk of the
    detector's radius, perpendicular to its direction; compared with the reference source on the detector's volume"""
    cfg = dict(benchmarks.get("cube60", 500000), issavedet=0, srcpos=[[30, 30, 1, 1], [30, 42, 1, 1]], srcdir=[[0, 0, 1, 0], [0, 0, 1, 0]],
               srcparam1=[[0, 0, 0, 0], [4, 0, 0, 0]], detpos=[[30, 42, 1, 4]], srcid=-1, outputtype="adjoint", isnormalized=0)
    p = hostcfg.prepare(cfg)
    assert p.c.extrasrclen == 1 and p.c.detnum == 1 and p.nsrcvol == 2
    g = engine.run_prepared(p)["field"].astype(np.float64).reshape(2, 60, 60, 60)
    o = ref.run(hostcfg.prepare(dict(cfg, nphoton=200000)), 2048, hostthreads=0)["field"].astype(np.float64).reshape(2, 60, 60, 60)
    g, o = g / 5e5, o / 2e5
    for k in (0, 1):
        assert g[k].sum() == pytest.approx(o[k].sum(), rel=0.02)
    # the entry layer of the detector volume: a disk of radius 4 voxels, not a point
    lg, lo = g[1, 0], o[1, 0]
    yy, xx = np.mgrid[0:60, 0:60]
    disk = (xx + 0.5 - 29) ** 2 + (yy + 0.5 - 41) ** 2 < 16
    assert lg[disk].sum() / lg.sum() == pytest.approx(lo[disk].sum() / lo.sum(), abs=0.02)
    assert lg[41, 29] / lg.sum() == pytest.approx(lo[41, 29] / lo.sum(), rel=0.2)
    # while the true source stays a pencil beam: its entry voxel holds several times the share the disk's centre voxel holds
    assert g[0, 0, 29, 29] / g[0, 0].sum() == pytest.approx(o[0, 0, 29, 29] / o[0, 0].sum(), rel=0.1)
    assert g[0, 0, 29, 29] / g[0, 0].sum() > 2.5 * lg[41, 29] / lg.sum()
