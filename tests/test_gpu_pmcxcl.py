"""The reference's UNCHANGED Python front-end (src/pmcxcl.cpp, pybind11 module `_pmcxcl`), relinked against the B200
engine through integration/mcx_cuda_host.cpp (integration/build_cli.py::build_pmcxcl).  The calls are the ones of the
reference's own documentation and benchmark table (pmcxcl/pmcxcl/__init__.py:19-30, pmcxcl/pmcxcl/bench.py:22-60);
the known answers are the reference's statistical pins (mcxlabcl/examples/mcx_gpu_benchmarks.m:65-112,
test/testmcx.sh:60-75)."""
import glob
import os
import sys

import numpy as np
import pytest

from mcxcl_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "integration", "_build")


def load_module():
    if not glob.glob(os.path.join(BUILD, "_pmcxcl*.so")):
        pytest.skip("integration/_build/_pmcxcl*.so not built (python integration/build_cli.py needs /root/reference)")
    if BUILD not in sys.path:
        sys.path.insert(0, BUILD)
    import _pmcxcl
    return _pmcxcl


def cube60(**kw):
    """pmcxcl/pmcxcl/bench.py:22-37"""
    cfg = dict(nphoton=1000000, vol=np.ones([60, 60, 60], dtype="uint8"), tstart=0, tend=5e-9, tstep=5e-9, srcpos=[29, 29, 0], srcdir=[0, 0, 1],
               prop=[[0, 0, 1, 1], [0.005, 1, 0.01, 1.37], [0.002, 5, 0.9, 1]], isreflect=0, seed=1648335518, session="cube60",
               detpos=[[29, 19, 0, 1], [29, 39, 0, 1], [19, 29, 0, 1], [39, 29, 0, 1]], issrcfrom0=1)
    cfg.update(kw)
    return cfg


def test_module_exports_the_reference_api():
    """runs without a GPU: the module imports and has run / gpuinfo / version (src/pmcxcl.cpp:1648-1663)"""
    m = load_module()
    assert callable(m.run) and callable(m.gpuinfo) and callable(m.version)
    assert m.version().startswith("v20")


@pytest.mark.gpu
def test_gpuinfo_lists_the_b200():
    m = load_module()
    info = m.gpuinfo()
    assert len(info) >= 1 and "B200" in info[0]["name"]
    mine = engine.gpuinfo()[0]
    assert info[0]["sm"] == mine["sm"] == 148 and info[0]["autothread"] == mine["autothread"]


@pytest.mark.gpu
@pytest.mark.parametrize("reflect,want", [(0, 0.1769), (1, 0.2701)])
def test_run_cube60_known_answers(reflect, want):
    """mcx_gpu_benchmarks.m:65-66, 87-88: absorbed fraction within 0.005, energytot within 10 of nphoton"""
    m = load_module()
    res = m.run(cube60(isreflect=reflect))
    st = res["stat"]
    assert abs(st["energytot"] - 1e6) <= 10
    assert abs(st["energyabs"] / st["energytot"] - want) < 0.005
    flux = res["flux"]
    assert flux.shape == (60, 60, 60, 1) and flux.dtype == np.float32 and np.isfinite(flux).all()
    # energy deposited = sum(flux * mua) * dt / normaliser-per-photon: the fluence integral reproduces the absorbed fraction
    absorbed = float((flux.astype(np.float64) * 0.005).sum() * 5e-9)
    assert abs(absorbed - st["energyabs"] / st["energytot"]) < 2e-3
    detp = res["detp"]
    assert detp.shape[0] == 1 + 2 and detp.shape[1] > 2000          # detid + partial path per medium (default savedetflag DP)
    assert set(np.unique(detp[0]).astype(int)) <= {1, 2, 3, 4}
    assert st["runtime"] > 0 and st["nphoton"] == 1000000 and st["unitinmm"] == 1.0


@pytest.mark.gpu
def test_run_keyword_form_and_planar_source():
    """pmcxcl.run(**cfg) (the README call) and a wide-field source: bench.py 'cube60planar', pin 25.x % (test/testmcx.sh:76-78)"""
    m = load_module()
    res = m.run(**cube60(isreflect=1, srctype="planar", srcpos=[10, 10, -10], srcparam1=[40, 0, 0, 0], srcparam2=[0, 40, 0, 0], nphoton=200000))
    st = res["stat"]
    assert 0.25 <= st["energyabs"] / st["energytot"] < 0.26


@pytest.mark.gpu
def test_run_matches_the_ctypes_mirror():
    """the same dictionary through the unchanged pybind11 front-end and through mcxcl_b200.engine.run: same RNG streams,
    same photon count => the normalised volumes agree to float rounding of the accumulation order"""
    m = load_module()
    cfg = cube60(isreflect=1, nphoton=100000)
    a = m.run(cfg)
    b = engine.run(dict(cfg))
    assert a["flux"].shape == b["flux"].shape
    # two realisations of 1e5 packets: sigma of the absorbed fraction is about 0.8 % of its value (tests/util.py:absorbed_sigma)
    assert abs(a["stat"]["energyabs"] - b["stat"]["energyabs"]) / b["stat"]["energyabs"] < 4e-2
    fa, fb = a["flux"].astype(np.float64), b["flux"].astype(np.float64)
    big = fb > 1e-3 * fb.max()
    # dynamic photon scheduling: streams are identical, their assignment to photons is not => statistical agreement only
    assert abs(fa[big].sum() / fb[big].sum() - 1) < 4e-2


@pytest.mark.gpu
def test_errors_surface_as_python_exceptions():
    """mcx_error -> mcx_throw_exception -> RuntimeError under MCX_CONTAINER (src/pmcxcl.cpp:1565-1568)"""
    m = load_module()
    with pytest.raises(RuntimeError, match="optical properties"):
        m.run(cube60(vol=3 * np.ones([60, 60, 60], dtype="uint8")))      # label 3 with a 3-row media table
    # a device that does not exist is not an error in the reference: it prints a notice and returns nothing
    # (src/pmcxcl.cpp:1147-1152)
    assert len(m.run(cube60(gpuid=9))) == 0


@pytest.mark.gpu
def test_replay_through_the_unchanged_front_end():
    """baseline with issaveseed -> res['seeds'] (16 x N bytes) and res['detp']; then cfg['seed'] = seeds, cfg['detphotons']
    = detp, outputtype 'jacobian' (src/pmcxcl.cpp:1006-1048; mcx_replayinit src/mcx_utils.c:1355-1421)"""
    m = load_module()
    base = m.run(cube60(issaveseed=1, savedetflag="dp", nphoton=300000))
    seeds, detp = base["seeds"], base["detp"]
    n = detp.shape[1]
    assert seeds.shape == (16, n) and n > 700
    rep = m.run(cube60(seed=seeds, detphotons=detp, savedetflag="dp", nphoton=300000))
    nrep = int(round(rep["stat"]["energytot"]))                        # mcx_replayinit may drop a record at the edge of the time window
    assert n - 3 <= nrep <= n and rep["detp"].shape[1] == nrep         # every replayed photon is detected again
    jac = m.run(cube60(seed=seeds, detphotons=detp, savedetflag="dp", outputtype="jacobian", nphoton=300000))["flux"]
    mine = engine.run(cube60(seed=seeds, detphotons=detp, savedetflag="dp", outputtype="jacobian", nphoton=300000))["flux"]
    assert jac.shape == mine.shape == (60, 60, 60, 1)
    # same records, same streams, same kernel: only the order of the floating-point accumulation differs
    np.testing.assert_allclose(jac.astype(np.float64).sum(), mine.astype(np.float64).sum(), rtol=1e-5)
    w = np.exp(-0.005 * detp[1] - 0.002 * detp[2])
    np.testing.assert_allclose(jac.astype(np.float64).sum(), float((w * (detp[1] + detp[2])).sum() / w.sum()), rtol=2e-3)


@pytest.mark.gpu
def test_rf_forward_through_the_unchanged_front_end():
    """cfg['omega'] > 0 (src/pmcxcl.cpp:469): pmcxcl allocates two volume sets and returns a complex flux (:1209-1211,
    1339-1400); the phase lag grows with depth and a vanishing frequency gives back the real run"""
    m = load_module()
    om = 2 * np.pi * 100e6
    res = m.run(cube60(isreflect=1, issavedet=0, nphoton=500000, omega=om))
    flux = res["flux"]
    assert np.iscomplexobj(flux) and flux.shape[:3] == (60, 60, 60)
    mine = engine.run(cube60(isreflect=1, issavedet=0, nphoton=500000, omega=om))["flux"]
    # two realisations of 5e5 packets: compared on a 5 x 5 column around the beam axis
    axis_a, axis_b = flux.reshape(60, 60, 60, -1)[27:32, 27:32, 1:22, 0].sum(axis=(0, 1)), mine[27:32, 27:32, 1:22, 0].sum(axis=(0, 1))
    np.testing.assert_allclose(np.abs(axis_a), np.abs(axis_b), rtol=0.1)
    assert np.max(np.abs(np.angle(axis_a / axis_b))) < 0.05
    ph = np.unwrap(np.angle(axis_a))
    assert ph[0] < 0 and ph[-1] < ph[0] - 0.2                           # lagging, and more so deeper in
    assert abs(res["stat"]["energyabs"] / res["stat"]["energytot"] - 0.2701) < 0.005


@pytest.mark.gpu
def test_adjoint_jacobian_through_the_unchanged_front_end():
    """outputtype 'adjoint' (src/pmcxcl.cpp:854-872, 1179-1198, 1414-1490): detectors are appended as sources, one fluence
    volume per source / detector, and 'jmua' = -Vvox x phi_src x phi_det per voxel (src/mcx_host.cpp:1468-1641)"""
    m = load_module()
    cfg = cube60(isreflect=0, nphoton=300000, detpos=[[29, 39, 0, 1]], detdir=[[0, 0, 1, 0]], outputtype="adjoint", issavedet=0)
    res = m.run(cfg)
    assert "jmua" in res and res["jmua"].shape == (60, 60, 60, 1)
    flux = np.asarray(res["flux"], dtype=np.float64).reshape((60, 60, 60, -1), order="F")
    assert flux.shape[-1] == 2                                          # {source, detector-as-source}
    want = -flux[..., 0] * flux[..., 1]                                 # Vvox = 1 mm^3, one gate
    got = res["jmua"][..., 0].astype(np.float64)
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-12)
    assert (got < 0).sum() > 10000                                      # a banana between the two optodes, not an empty volume
    both = m.run(dict(cfg, outputtype="adjoint_mua_d"))
    assert "jmua" in both and "jd" in both and both["jd"].shape == (60, 60, 60, 1)
    vols = np.asarray(both["flux"], dtype=np.float64).reshape((60, 60, 60, -1), order="F")
    g0, g1 = np.gradient(vols[..., 0], edge_order=2), np.gradient(vols[..., 1], edge_order=2)
    dot = sum(a * b for a, b in zip(g0, g1))                            # second-order differences, like mcx_fd_grad
    np.testing.assert_allclose(both["jd"][..., 0].astype(np.float64), -dot, rtol=2e-3, atol=1e-6 * np.abs(dot).max())


@pytest.mark.gpu
def test_polarised_run_through_the_unchanged_front_end():
    """cfg['polprop'] + cfg['lambda'] (src/pmcxcl.cpp:754-800, 470): the front-end's own Mie code fills prop and the Mueller
    matrices (mcx_prep_polarized, src/mcx_utils.c:1483-1519); savedetflag 'i' appends the Stokes vector (src/mcx_core.cl:912-917)"""
    m = load_module()
    cfg = cube60(isreflect=1, nphoton=500000, prop=[[0, 0, 1, 1], [0.005, 1, 0.01, 1.37]], polprop=[[0.005, 0.05, 19.11, 1.59, 1.33]],
                 savedetflag="dpi", srciquv=[1, 1, 0, 0])
    cfg["lambda"] = 632.8
    res = m.run(cfg)
    detp = res["detp"]
    assert detp.shape[0] == 1 + 1 + 3 + 1 + 4 and detp.shape[1] > 300   # D, P, and -- forced for polarised runs (src/mcx_utils.c:1777-1781) -- V, W, I
    s = detp[-4:]
    assert np.all(s[0] == 1.0) and np.all(np.abs(s[1:]) <= 1.0 + 1e-4) and np.abs(s[1:3]).mean() > 0.05
    st = res["stat"]
    assert 0.05 < st["energyabs"] / st["energytot"] < 0.9 and abs(st["energytot"] - 5e5) <= 10
