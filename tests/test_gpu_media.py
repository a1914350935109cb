"""Continuous media on the B200: volumes whose 32-bit words ENCODE the optical properties of every voxel (Config.mediabyte
99-104, decoded per segment like updateproperty, src/mcx_core.cl:1079-1193), against the reference kernel source built
with the same -DMED_TYPE (oracle/_ref).  The words are packed like pmcxcl packs them (src/pmcxcl.cpp:108-400)."""
import numpy as np
import pytest

import decks
from mcxcl_b200 import benchmarks, engine, hostcfg
from util import absorbed_sigma, run_gpu, run_ref

pytestmark = pytest.mark.gpu
N = 200000
PROP2, PROP3, two_regions, volumes = decks.MEDIA_PROP2, decks.MEDIA_PROP3, decks.media_two_regions, decks.media_volumes


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
