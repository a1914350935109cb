"""Continuous media on the B200: volumes whose 32-bit words ENCODE the optical properties of every voxel (Config.mediabyte
99-104, decoded per segment like updateproperty, src/mcx_core.cl:1079-1193), against the reference kernel source built
with the same -DMED_TYPE (oracle/_ref).  The words are packed like pmcxcl packs them (src/pmcxcl.cpp:108-400)."""
import numpy as np
import pytest

from mcxcl_b200 import benchmarks, engine, hostcfg
from util import absorbed_sigma, run_gpu, run_ref

pytestmark = pytest.mark.gpu
N = 200000
PROP2 = [[0, 0, 1, 1], [0.005, 1.0, 0.01, 1.37]]
PROP3 = [[0, 0, 1, 1], [0.0, 0.0, 0.01, 1.0], [0.02, 5.0, 0.9, 1.5]]          # rows 1 and 2 = the range of the scaled formats


def two_regions():
    """mua / mus maps: a 60^3 cube with a more absorbing, more scattering slab at z = 20..39"""
    mua = np.full((60, 60, 60), 0.005, np.float32)
    mus = np.full((60, 60, 60), 1.0, np.float32)
    mua[:, :, 20:40] = 0.015
    mus[:, :, 20:40] = 2.5
    return mua, mus


def volumes():
    mua, mus = two_regions()
    out = {}
    out["mua_float"] = (mua[None], PROP2, hostcfg.MEDIA_MUA_FLOAT)
    out["as_f2h"] = (np.stack([mua, mus]), PROP2, hostcfg.MEDIA_AS_F2H)
    lh = np.zeros((3, 60, 60, 60), np.float32)            # {value, slot, label}: mua of label 1 replaced inside the slab
    lh[0], lh[1], lh[2] = mua, 0, 1
    out["label_half"] = (lh, PROP2, hostcfg.MEDIA_LABEL_HALF)
    b = np.zeros((4, 60, 60, 60), np.uint8)                # bytes scale between PROP3 rows 1 and 2
    b[0] = np.round(mua / 0.02 * 255)
    b[1] = np.round(mus / 5.0 * 255)
    b[2] = 0
    b[3] = 94                                              # n = 1 + 94/127 * 0.5 = 1.37
    b[3, :, :, 20:40] = 127                                # ... and 1.5 inside the slab: interior Fresnel faces
    out["asgn_byte"] = (b, PROP3, hostcfg.MEDIA_ASGN_BYTE)
    sh = np.zeros((2, 60, 60, 60), np.uint16)
    sh[0] = np.round(mua / 0.02 * 65535)
    sh[1] = np.round(mus / 5.0 * 65535)
    out["as_short"] = (sh, [[0, 0, 1, 1], [0.0, 0.0, 0.01, 1.37], [0.02, 5.0, 0.01, 1.37]], hostcfg.MEDIA_AS_SHORT)
    return out


@pytest.mark.parametrize("name", ["mua_float", "as_f2h", "label_half", "asgn_byte", "as_short"])
@pytest.mark.parametrize("reflect", [0, 1])
def test_continuous_media_match_reference(ref, name, reflect):
    vol, prop, fmt = volumes()[name]
    cfg = dict(benchmarks.get("cube60b", N), vol=vol, prop=prop, isreflect=reflect, issavedet=0)
    p = hostcfg.prepare(cfg)
    assert p.c.mediaformat == fmt
    with engine.Simulation(p) as sim:
        assert "uint32_t" in sim.kernel_name and sim.kernel_name.endswith("/true/q0")       # 32-bit media words, generic kernel
    p, r = run_gpu(cfg)
    _, o = run_ref(ref, cfg, work=1024)
    assert r["energytot"] == N
    sig = np.hypot(absorbed_sigma(N, r["absorbed"]), absorbed_sigma(N, o["absorbed"]))
    assert abs(r["absorbed"] - o["absorbed"]) < 5 * sig, (r["absorbed"], o["absorbed"])
    gf = (r["field"].astype(np.float64) / r["normalizer"]).reshape(60, 60, 60)             # [z][y][x]
    of = o["field"].astype(np.float64).reshape(60, 60, 60)
    for z0, z1, tol in ((0, 20, 0.02), (20, 40, 0.03), (40, 60, 0.08)):
        np.testing.assert_allclose(gf[z0:z1].sum(), of[z0:z1].sum(), rtol=tol)
    np.testing.assert_allclose(gf.sum(axis=(1, 2))[:30], of.sum(axis=(1, 2))[:30], rtol=0.08)


def test_continuous_media_equal_the_label_run_they_encode():
    """a MUA_FLOAT volume that holds the mua of label 1 everywhere is the cube60b benchmark itself"""
    base = dict(benchmarks.get("cube60b", N), issavedet=0)
    _, a = run_gpu(base)
    _, b = run_gpu(dict(base, vol=np.full((1, 60, 60, 60), 0.005, np.float32)))
    sig = absorbed_sigma(N, a["absorbed"])
    assert abs(a["absorbed"] - b["absorbed"]) < 5 * np.sqrt(2.0) * sig
    np.testing.assert_allclose(b["field"].astype(np.float64).sum(), a["field"].astype(np.float64).sum(), rtol=0.02)     # 2e5 packets each: sigma of the ratio 0.4 %


def test_continuous_media_with_detectors_and_refusals():
    mua, mus = two_regions()
    cfg = dict(benchmarks.get("cube60b", N), vol=np.stack([mua, mus]), prop=PROP2)
    # detector id, exit position / direction and launch weight need no per-medium rows: allowed
    p, r = run_gpu(cfg, savedetflag="dxvw")
    assert r["reclen"] == 8 and r["detected"] > 300 and set(np.unique(r["detp"][:, 0]).astype(int)) == {1, 2, 3, 4}
    # the per-medium columns would be indexed with the media word (the reference does exactly that, src/mcx_core.cl:2515)
    with pytest.raises(RuntimeError, match="label media"):
        run_gpu(cfg)                                        # default record "DP"
    with pytest.raises(RuntimeError, match="outside this build"):
        run_gpu(dict(cfg, vol=(np.ones((60, 60, 60), np.uint32)), mediaformat=96))     # two-word media: no front-end produces them
    with pytest.raises(hostcfg.ConfigError):
        hostcfg.prepare(dict(cfg, prop=[[0, 0, 1, 1]]))
