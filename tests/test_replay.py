"""Photon replay: saved RNG states of detected photons are re-launched and every one of
them follows its original trajectory.  Reference: kernel src/mcx_core.cl:1590-1596 (stream restart), :2567-2592
(scattering-site outputs), :2845-2858 (absorption Jacobian); host src/mcx_host.cpp:684-689, 722-737, 1398-1432; set-up
src/mcx_utils.c:1355-1421 (mcx_replayinit).

Pins the reference's own tests hold for this path (test/testmcx.sh:116-128): the replay detects exactly as many photons
as it launches, and cube60 replays absorb 3[0-8].x %.  Everything else is pinned by running the reference's kernel
source itself (oracle/_ref) and by identities that hold photon by photon:
    sum over voxels of the raw absorption Jacobian of medium m   ==  sum_i  w_i * L_i(m)
    sum over voxels of the raw WP output                        ==  sum_i  w_i * (number of scattering events of i)
    sum over voxels of the raw DCS output                       ==  sum_i  w_i * (momentum transfer of i)
with w_i, L_i, nscat_i, mom_i taken from the baseline run's detected-photon records."""
import numpy as np
import pytest

from mcxcl_b200 import benchmarks, engine, hostcfg


def baseline_cfg(nphoton=100000, **kw):
    cfg = benchmarks.get("cube60", nphoton)
    cfg.update(issaveseed=1, savedetflag="DSPM")
    cfg.update(kw)
    return cfg


def replay_cfg(cfg, detp, seeds, **kw):
    """what a pmcxcl user passes (src/pmcxcl.cpp:1006-1048): cfg['seed'] = uint8[16, N], cfg['detphotons'] = detp"""
    rc = dict(cfg)
    rc.pop("issaveseed", None)
    rc.update(seed=np.ascontiguousarray(seeds).view(np.uint8).reshape(-1, 16).T.copy(), detphotons=np.ascontiguousarray(detp.T))
    rc.update(kw)
    return rc


def sort_rows(a):
    return a[np.lexsort(a.T[::-1])]


def expected_sums(p, detp):
    """per-photon records -> the right-hand sides of the identities in the module docstring (media 1..M)"""
    M = p.c.medianum - 1
    w = p.keep["replay_weight"].astype(np.float64)
    detp = detp[p.keep["replay_index"]]         # mcx_replayinit drops records outside the time window (src/mcx_utils.c:1408-1411)
    # the scattering counts are uint32 bit patterns inside the float record (src/mcx_core.cl:2503-2520)
    nscat = np.ascontiguousarray(detp[:, 1:1 + M]).view(np.uint32).astype(np.float64)
    plen, mom = detp[:, 1 + M:1 + 2 * M], detp[:, 1 + 2 * M:1 + 3 * M]
    return (w[:, None] * plen).sum(), (w[:, None] * nscat).sum(), (w[:, None] * mom).sum()


def shift_records_for_the_reference(p):
    """The reference's OpenCL kernel indexes replayweight / photontof / photondetid with `... + (int)f.w` at scattering
    sites (src/mcx_core.cl:2569-2586) but with `... + (int)f.w - 1` in the Jacobian branch (:2847-2849); f.w was already
    incremented by the launch, so the WP / DCS / WPTOF outputs of photon i are weighted with the record of photon i+1
    (and the last photon reads one element past the buffers).  This engine pairs photon i with record i.  To check the
    CUDA kernel against the reference SOURCE for these outputs, hand the reference tables that are padded by one leading
    element, so that its i+1 read lands on record i."""
    import ctypes as C
    for key, ctype in (("replay_weight", C.c_float), ("replay_tof", C.c_float), ("replay_detid", C.c_int32)):
        a = p.keep[key]
        padded = np.concatenate([a[:1], a]).astype(a.dtype)
        p.keep[key + "_padded"] = padded
        setattr(p.c, key, padded.ctypes.data_as(C.POINTER(ctype)))
    return p


# ------------------------------------------------------------------------------------------- CPU: the oracle, pinned
def test_oracle_replay_reproduces_the_reference_pins(ref):
    cfg = baseline_cfg()
    base = ref.run(hostcfg.prepare(cfg), 1024, hostthreads=0)
    n = base["detected"]
    assert 200 < n < 400 and base["seeds"].shape == (n, 2)
    pr = hostcfg.prepare(replay_cfg(cfg, base["detp"], base["seeds"]))
    assert pr.c.nphoton == n and pr.c.seed == hostcfg.SEED_FROM_FILE
    rep = ref.run(pr, 1, hostthreads=0)
    assert rep["detected"] == n                                         # test/testmcx.sh:116-121
    assert 0.30 <= rep["absorbed"] < 0.39                               # test/testmcx.sh:123-128
    assert np.array_equal(sort_rows(base["detp"]), sort_rows(rep["detp"]))   # same trajectories, bit for bit


def test_replayinit_mirror_matches_the_formula():
    """mcx_replayinit (src/mcx_utils.c:1397-1411): w = prod exp(-mua_m L_m), tof = sum L_m unitinmm n_m / c0; photons of
    another detector or outside the time window are dropped"""
    cfg = baseline_cfg(10, savedetflag="DP")
    detp = np.array([[1, 10.0, 0.0], [2, 20.0, 0.0], [1, 1e6, 0.0], [1, 5.0, 2.0]], np.float32)
    seeds = np.arange(8, dtype=np.uint64).reshape(4, 2)
    p = hostcfg.prepare(replay_cfg(cfg, detp, seeds, replaydet=1))
    assert p.c.nphoton == 2 and p.keep["replay_detid"].tolist() == [1, 1]
    np.testing.assert_allclose(p.keep["replay_weight"], [np.exp(-0.05), np.exp(-0.025 - 0.004)], rtol=1e-6)
    np.testing.assert_allclose(p.keep["replay_tof"], [10 * 1.37 * 3.335640951981520e-12, (5 * 1.37 + 2 * 1.0) * 3.335640951981520e-12], rtol=1e-6)
    assert p.keep["replay_seed"].tolist() == [[0, 1], [6, 7]]
    with pytest.raises(hostcfg.ConfigError):
        hostcfg.prepare(dict(baseline_cfg(10), outputtype="jacobian"))     # sensitivity outputs exist only in replay


@pytest.mark.parametrize("ot,which", [("jacobian", 0), ("wp", 1), ("wm", 2)])
def test_oracle_sensitivity_outputs_satisfy_the_identities(ref, ot, which):
    cfg = baseline_cfg()
    base = ref.run(hostcfg.prepare(cfg), 1024, hostthreads=0)
    p = hostcfg.prepare(replay_cfg(cfg, base["detp"], base["seeds"], outputtype=ot, isnormalized=0))
    want = expected_sums(p, base["detp"])[which]
    if which:
        shift_records_for_the_reference(p)          # see the helper: a reference off-by-one in the scattering-site outputs
    out = ref.run(p, 1, hostthreads=0)
    assert out["field"].astype(np.float64).sum() == pytest.approx(want, rel=2e-5)


# ------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def gpu_base():
    cfg = baseline_cfg(1000000)
    r = engine.run_prepared(hostcfg.prepare(cfg))
    assert r["saved"] == r["detected"] > 2500
    return cfg, r


@pytest.mark.gpu
def test_gpu_replay_retraces_every_photon(gpu_base):
    cfg, base = gpu_base
    p = hostcfg.prepare(replay_cfg(cfg, base["detp"], base["seeds"], issaveseed=1))
    rep = engine.run_prepared(p)
    kept = p.keep["replay_index"]               # all but the odd record whose time of flight rounds past tend
    assert base["detected"] - 3 <= p.c.nphoton == len(kept) <= base["detected"]
    assert rep["energytot"] == p.c.nphoton
    assert rep["detected"] == p.c.nphoton
    # pair the records through the saved RNG states (unique per packet).  The baseline runs in a common-configuration
    # kernel and the replay in the generic one: same arithmetic, but the two instantiations may contract a*b+c
    # differently, so integer fields must agree exactly and float fields to rounding
    bs, rs = base["seeds"][kept].astype(np.int64), rep["seeds"].astype(np.int64)
    ob, orr = np.lexsort(bs.T[::-1]), np.lexsort(rs.T[::-1])
    assert np.array_equal(bs[ob], rs[orr])
    b, r = base["detp"][kept][ob], rep["detp"][orr]
    M = p.c.medianum - 1
    assert np.array_equal(b[:, 0], r[:, 0])                                             # detector id
    assert np.array_equal(b[:, 1:1 + M].view(np.uint32), r[:, 1:1 + M].view(np.uint32))  # scattering counts
    np.testing.assert_allclose(b[:, 1 + M:], r[:, 1 + M:], rtol=2e-4, atol=1e-5)         # partial paths, momentum transfer
    assert 0.30 <= rep["absorbed"] < 0.39


@pytest.mark.gpu
@pytest.mark.parametrize("ot,which", [("jacobian", 0), ("wp", 1), ("wm", 2)])
def test_gpu_sensitivity_identities(gpu_base, ot, which):
    cfg, base = gpu_base
    p = hostcfg.prepare(replay_cfg(cfg, base["detp"], base["seeds"], outputtype=ot, isnormalized=0))
    out = engine.run_prepared(p)
    want = expected_sums(p, base["detp"])[which]
    assert out["field"].astype(np.float64).sum() == pytest.approx(want, rel=2e-5)
    # normalised form: unitinmm / sum of the detected weights (src/mcx_host.cpp:1424-1432)
    pn = hostcfg.prepare(replay_cfg(cfg, base["detp"], base["seeds"], outputtype=ot))
    outn = engine.run_prepared(pn)
    scale = 1.0 / float(p.keep["replay_weight"].astype(np.float32).sum())
    assert outn["normalizer"] == pytest.approx(scale, rel=1e-5)
    np.testing.assert_allclose(outn["field"].astype(np.float64).sum(), want * scale, rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("ot", ["jacobian", "wp", "wm"])
def test_gpu_sensitivities_match_the_reference_source_voxel_by_voxel(gpu_base, ref, ot):
    """the SAME records replayed by the reference's kernel source on the host and by the CUDA kernel: the two walk the
    same RNG streams; exp/log/sincos differ in the last bits (fast tier), so a few long paths decorrelate -- the
    volumes agree closely but not bit for bit"""
    cfg, base = gpu_base
    n = 1500
    rc = replay_cfg(cfg, base["detp"][:n], base["seeds"][:n], outputtype=ot, isnormalized=0)
    g = engine.run_prepared(hostcfg.prepare(rc))["field"].astype(np.float64)
    po = hostcfg.prepare(rc)
    if ot != "jacobian":
        shift_records_for_the_reference(po)
    o = ref.run(po, 1, hostthreads=0)["field"].astype(np.float64)
    assert g.sum() == pytest.approx(o.sum(), rel=2e-3)
    assert np.corrcoef(g, o)[0, 1] > 0.98
    hot = o > 0.05 * o.max()
    assert hot.sum() > 20 and np.abs(g[hot] / o[hot] - 1).mean() < 0.05


@pytest.mark.gpu
def test_gpu_replay_of_every_detector_at_once_and_time_gates(gpu_base):
    """replaydet = -1: one volume per detector (src/mcx_host.cpp:684-689), each normalised by the weights detected there
    (:1398-1421); the time gate of a deposit is the DETECTED time of flight of the photon (:2849-2851)"""
    cfg, base = gpu_base
    rc = replay_cfg(cfg, base["detp"], base["seeds"], outputtype="jacobian", isnormalized=0, replaydet=-1, tstep=1e-9)
    p = hostcfg.prepare(rc)
    assert p.maxgate == 5 and p.nrepvol == 4 and p.fieldlen == 216000 * 5 * 4
    out = engine.run_prepared(p)
    vol = out["field"].astype(np.float64).reshape(4, 5, 216000)
    w, tof, det = p.keep["replay_weight"].astype(np.float64), p.keep["replay_tof"], p.keep["replay_detid"]
    plen = base["detp"][p.keep["replay_index"], 3:5].sum(axis=1)
    gate = np.minimum(np.floor(tof / np.float32(1e-9)).astype(int), 4)
    for d in range(1, 5):
        for g in range(5):
            sel = (det == d) & (gate == g)
            assert vol[d - 1, g].sum() == pytest.approx((w[sel] * plen[sel]).sum(), rel=5e-5, abs=1e-6)
    flux = engine.run(dict(rc, isnormalized=1))["flux"]
    assert flux.shape == (60, 60, 60, 5, 4)
    for d in range(1, 5):
        np.testing.assert_allclose(flux[..., d - 1].astype(np.float64).sum(), vol[d - 1].sum() / w[det == d].astype(np.float32).sum(), rtol=1e-4)


@pytest.mark.gpu
def test_gpu_replay_refuses_what_it_cannot_do(gpu_base):
    cfg, base = gpu_base
    p = hostcfg.prepare(replay_cfg(cfg, base["detp"], base["seeds"], outputtype="jacobian"))
    p.c.seed_skip = 5                        # a second device's slice: replay is single-device (src/mcx_host.cpp:723)
    with pytest.raises(RuntimeError, match="single device"):
        engine.run_prepared(p)


@pytest.mark.gpu
def test_progress_poll_while_the_kernel_runs():
    """mcxb_sim_progress: what `-D P` shows (src/mcx_host.cpp:1112-1141)"""
    import time
    p = hostcfg.prepare(benchmarks.get("cube60", 30000000))
    with engine.Simulation(p) as sim:
        sim.reset()
        sim.launch()
        seen = []
        while True:
            n, done = sim.progress()
            seen.append(n)
            if done:
                break
            time.sleep(0.005)
        assert seen == sorted(seen) and seen[-1] == 30000000
        assert any(0 < s < 30000000 for s in seen)          # observed mid-flight, i.e. the poll did not wait for the kernel
        assert sim.fetch()["energytot"] == 30000000
