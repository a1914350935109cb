"""Photon sharing: one packet of unit weight carries `srcnum` illumination patterns; each
deposit is scaled by the pattern values of the launch cell and goes to that pattern's volume.  Reference: launch
src/mcx_core.cl:1694-1705, deposit :2902-2911 (volumes interleaved pattern-fastest), host post-processing
src/mcx_host.cpp:1351-1380 (per-pattern totals) and :1436-1462 (scale_i = psize / sum(pattern_i) x scale)."""
import numpy as np
import pytest

from mcxcl_b200 import benchmarks, engine, hostcfg

NX = NY = 8


def patterns():
    """three patterns on an 8x8 aperture: uniform, the same x2 (exactly proportional), and a half-plane mask"""
    a = np.ones((NX, NY), np.float32)
    c = np.zeros((NX, NY), np.float32)
    c[:NX // 2] = 1.0
    return np.stack([a, 2 * a, c])            # (srcnum, Nx, Ny): what pmcxcl takes (README.md, 'srcpattern')


def sharing_cfg(nphoton, pat=None, **kw):
    cfg = benchmarks.get("cube60planar", nphoton)
    pat = patterns() if pat is None else pat
    cfg.update(srctype="pattern", srcparam1=[40.0, 0.0, 0.0, NX], srcparam2=[0.0, 40.0, 0.0, NY], srcpattern=pat,
               srcnum=pat.shape[0] if pat.ndim == 3 else 1, issavedet=0)
    cfg.update(kw)
    return cfg


def volumes(p, field):
    """raw float32[fieldlen] -> (srcnum, dimxyz*maxgate)"""
    return field.reshape(-1, p.c.srcnum).T


def test_prepare_sizes_the_output_per_pattern():
    p = hostcfg.prepare(sharing_cfg(10))
    assert p.c.srcnum == 3 and p.nsrcvol == 3 and p.fieldlen == 3 * 216000
    # pattern-fastest memory order: element [cell*srcnum + i] is pattern i at that cell
    pat = p.keep["srcpattern"].reshape(-1, 3)
    assert np.array_equal(pat[:, 1], 2 * pat[:, 0]) and pat[:, 2].sum() == NX * NY / 2
    with pytest.raises(hostcfg.ConfigError):
        hostcfg.prepare(sharing_cfg(10, srctype="planar"))


def test_oracle_photon_sharing_matches_one_pattern_at_a_time(ref):
    """the reference source itself: the half-plane volume of a shared run agrees statistically with a run of that
    pattern alone (same estimator, independent packets)"""
    n = 60000
    p = hostcfg.prepare(sharing_cfg(n, isnormalized=0))
    shared = volumes(p, ref.run(p, 1024, hostthreads=0)["field"].astype(np.float64))
    assert np.array_equal(shared[1], 2 * shared[0])
    p1 = hostcfg.prepare(sharing_cfg(n, pat=patterns()[2], isnormalized=0, seed=12345))
    alone = ref.run(p1, 1024, hostthreads=0)["field"].astype(np.float64)
    # run alone, zero-weight cells are re-drawn (:2090-2100), so all n packets start in the lit half; shared, half of them do
    assert shared[2].sum() == pytest.approx(alone.sum() / 2, rel=0.02)


@pytest.mark.gpu
def test_gpu_shared_volumes_are_exact_multiples():
    """patterns 0 and 1 differ by an exact factor 2: the same packets deposit into both, so the raw volumes are
    bit-exact multiples; normalised by their own totals (psize / sum(pattern)) they become equal"""
    p = hostcfg.prepare(sharing_cfg(500000, isnormalized=0))
    r = engine.run_prepared(p)
    assert r["energytot"] == 500000                     # unit-weight packets (:1704)
    v = volumes(p, r["field"])
    assert v[0].sum() > 0 and np.array_equal(v[1], 2 * v[0])
    assert 0.45 < v[2].sum() / v[0].sum() < 0.55
    pn = hostcfg.prepare(sharing_cfg(500000))
    vn = volumes(pn, engine.run_prepared(pn)["field"])
    np.testing.assert_allclose(vn[1], vn[0], rtol=1e-6)
    # energy balance per pattern: sum(flux * mua) * Vvox * dt == absorbed fraction of THAT pattern's launched energy
    absorbed = [(vn[i].astype(np.float64) * 0.005).sum() * 5e-9 for i in range(3)]
    assert absorbed[0] == pytest.approx(r["absorbed"], rel=2e-3)
    assert 0.2 < absorbed[2] < 0.3
    flux = engine.run(sharing_cfg(1000))["flux"]
    assert flux.shape == (3 * 60, 60, 60, 1)


@pytest.mark.gpu
def test_gpu_photon_sharing_against_the_reference_source(ref):
    n = 200000
    p = hostcfg.prepare(sharing_cfg(n, isnormalized=0))
    g = volumes(p, engine.run_prepared(p)["field"].astype(np.float64))
    o = volumes(p, ref.run(hostcfg.prepare(sharing_cfg(60000, isnormalized=0)), 1024, hostthreads=0)["field"].astype(np.float64)) * (n / 60000)
    for i in range(3):
        assert g[i].sum() == pytest.approx(o[i].sum(), rel=0.02)       # the 6e4-packet reference run alone carries 0.5 % (0.7 % for the half-plane pattern)
        # depth profile of each pattern's volume (z is the slowest axis)
        gz, oz = g[i].reshape(60, -1).sum(axis=1), o[i].reshape(60, -1).sum(axis=1)
        np.testing.assert_allclose(gz[:30], oz[:30], rtol=0.06)
    # the half-plane pattern illuminates x < 30 only: its volume is lopsided, the uniform one is not
    gx = g[2].reshape(60, 60, 60).sum(axis=(0, 1))
    assert gx[:30].sum() > 3 * gx[30:].sum()
    ux = g[0].reshape(60, 60, 60).sum(axis=(0, 1))
    assert ux[:30].sum() == pytest.approx(ux[30:].sum(), rel=0.03)


@pytest.mark.gpu
def test_gpu_photon_sharing_time_gates_and_single_pattern_limit():
    """two gates: layout [gate][voxel][pattern]; a one-pattern 'sharing' deck is the ordinary pattern source"""
    p = hostcfg.prepare(sharing_cfg(200000, isnormalized=0, tstep=2.5e-9))
    r = engine.run_prepared(p)
    f = r["field"].astype(np.float64).reshape(2, 216000, 3)
    assert f[0].sum() > f[1].sum() > 0
    assert np.array_equal(f[..., 1], 2 * f[..., 0])
    one = hostcfg.prepare(sharing_cfg(200000, pat=patterns()[2], isnormalized=0, tstep=2.5e-9))
    f1 = engine.run_prepared(one)["field"].astype(np.float64).reshape(2, 216000)
    assert f[..., 2].sum() == pytest.approx(f1.sum() / 2, rel=0.02)        # alone, zero-weight cells are re-drawn: twice the packets in the lit half
