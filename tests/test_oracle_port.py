"""The plain-C restatement (oracle/mcx_oracle.c) against the reference's own kernel source (oracle/_ref) and against
the known answers of tests/golden/kat_survey.json.

Both checkers run under the same numeric contract (IEEE binary32, no contraction, libm float functions), so with
one host thread -- work-items executed in index order into one private field -- every output must agree BIT FOR BIT:
field, per-work-item energies, detected-photon records, saved seeds and the work counters."""
import json
import os

import numpy as np
import pytest

import decks
from mcxcl_b200 import benchmarks, hostcfg

HERE = os.path.dirname(os.path.abspath(__file__))
KAT = json.load(open(os.path.join(HERE, "golden", "kat_survey.json")))


def bits(x):
    return np.asarray(x, dtype=np.float32).view(np.uint32)


def both(ref, port, cfg, work=64):
    p = hostcfg.prepare(cfg)
    a = ref.run(p, work, hostthreads=1, want_energy=True)
    b = port.run(p, work, hostthreads=1, want_energy=True)
    return p, a, b


def assert_identical(a, b):
    assert a["energytot"] == b["energytot"] and a["energyesc"] == b["energyesc"]
    assert (bits(a["energy"]) == bits(b["energy"])).all()
    assert (bits(a["field"]) == bits(b["field"])).all()
    assert a["detected"] == b["detected"] and a["reclen"] == b["reclen"]
    if a["detp"] is not None:
        assert (bits(a["detp"]) == bits(b["detp"])).all()
    if a["seeds"] is not None:
        assert (a["seeds"] == b["seeds"]).all()
    assert (a["n_segment"], a["n_deposit"], a["n_scatter"]) == (b["n_segment"], b["n_deposit"], b["n_scatter"])


# ---------------------------------------------------------------------------------------- unit level

def test_port_rng_kat(port):
    seeds = port.seeds(KAT["seed"], len(KAT["rng"]))
    for row in KAT["rng"]:
        assert seeds.reshape(-1, 4)[row["thread"]].tolist() == row["seed"]
    u, _ = port.rng(seeds, 3)
    _, st = port.rng(seeds, 0)
    for row in KAT["rng"]:
        np.testing.assert_allclose(u[row["thread"]], np.array(row["u"], dtype=np.float32), rtol=2e-7, atol=0)
        assert [int(x) for x in st[row["thread"]]] == [int(t, 16) for t in row["t"]]


def test_port_traversal_and_scalar_kat(port):
    t = KAT["trace"]
    v0 = np.array([int(b, 16) for b in t["v0_bits"]], dtype=np.uint32).view(np.float32)
    out = port.trace(np.array(t["p0"] + [1.0], dtype=np.float32), np.append(v0, 0).astype(np.float32), len(t["steps"]), t["dims"], t["musp"])[0]
    for k, s in enumerate(t["steps"]):
        assert int(out["face"][k]) == s["face"] and int(out["idx1d"][k]) == s["idx1d"]
        assert "%08x" % bits(out["dist"][k]) == s["dist"]
        assert ["%08x" % bits(out[c][k]) for c in ("px", "py", "pz")] == s["p"]
    sc = KAT["scalar"]
    rc = sc["reflectcoeff"]
    na, r = port.scalar([x[0] for x in sc["nextafter"]], [x[1] for x in sc["nextafter"]], [x["v"] + [0] for x in rc],
                        [x["n1"] for x in rc], [x["n2"] for x in rc], [x["face"] for x in rc])
    assert ["%08x" % b for b in bits(na)] == [x[2] for x in sc["nextafter"]]
    assert ["%08x" % b for b in bits(r)] == [x["bits"] for x in rc]


def test_port_unit_hooks_match_reference_source(ref, port):
    rs = np.random.RandomState(5)
    seeds = rs.randint(0, 2 ** 31, 4 * 300).astype(np.uint32)
    ua, sa = ref.rng(seeds, 40)
    ub, sb = port.rng(seeds, 40)
    assert (bits(ua) == bits(ub)).all() and (sa == sb).all()
    n = 400
    p0 = np.zeros((n, 4), np.float32)
    p0[:, :3] = rs.uniform(0, 1, (n, 3)) * [31, 17, 23]
    v0 = np.zeros((n, 4), np.float32)
    d = rs.normal(size=(n, 3))
    d[:40, 0] = 0          # axis-parallel rays: the division by zero of hitgrid
    d[40:80, 1:] = 0
    v0[:, :3] = d / np.linalg.norm(d, axis=1, keepdims=True)
    ta = ref.trace(p0, v0, 48, (31, 17, 23), 2.5)
    tb = port.trace(p0, v0, 48, (31, 17, 23), 2.5)
    assert all(ta[f].tobytes() == tb[f].tobytes() for f in ta.dtype.names)
    a = rs.uniform(-5, 70, 500).astype(np.float32)
    dr = rs.randint(-1, 2, 500).astype(np.int32)
    n1 = rs.choice([1.0, 1.33, 1.37, 1.5], n).astype(np.float32)
    n2 = rs.choice([1.0, 1.33, 1.37, 1.5], n).astype(np.float32)
    face = rs.randint(0, 3, n).astype(np.int32)
    ra = ref.scalar(a, dr, v0, n1, n2, face)
    rb = port.scalar(a, dr, v0, n1, n2, face)
    assert (bits(ra[0]) == bits(rb[0])).all() and (bits(ra[1]) == bits(rb[1])).all()
    th, ph = rs.uniform(0, np.pi, n), rs.uniform(0, 2 * np.pi, n)
    v1 = v0.copy()
    v1[:10, :3] = [0, 0, 1]
    v1[10:20, :3] = [0, 0, -1]
    args = [np.sin(th), np.cos(th), np.sin(ph), np.cos(ph)]
    assert (bits(ref.rotate(v1, *args)) == bits(port.rotate(v1, *args))).all()
    ok = n1 <= n2          # transmission always defined
    assert (bits(ref.transmit(v0[ok], n1[ok], n2[ok], face[ok])) == bits(port.transmit(v0[ok], n1[ok], n2[ok], face[ok]))).all()


# ----------------------------------------------------------------------------------- whole simulations

@pytest.mark.parametrize("name", sorted(decks.SOURCES))
def test_port_bit_identical_sources(ref, port, name):
    cfg = decks.cube(3000, **decks.SOURCES[name])
    _, a, b = both(ref, port, cfg)
    assert a["energytot"] > 0
    assert_identical(a, b)


@pytest.mark.parametrize("name", sorted(decks.BOUNDARIES))
def test_port_bit_identical_boundaries(ref, port, name):
    over = dict(decks.BOUNDARIES[name])
    over["nphoton"] = min(int(over.get("nphoton", 3000)), 3000)
    _, a, b = both(ref, port, decks.cube(**over))
    assert_identical(a, b)


@pytest.mark.parametrize("case", ["two_layer", "two_layer_gates", "roulette", "savedet_all", "saveseed", "energy", "length",
                                  "multisrc_pick", "multisrc_volumes", "multisrc_fixed", "gscatter", "saveref", "flat2d",
                                  "skinvessel", "phase_table", "angle_table", "angle_table_discrete", "rngdebug", "odd_split"])
def test_port_bit_identical_features(ref, port, case):
    multi = dict(srcpos=[[29.0, 29.0, 0.0, 1.0], [10.0, 40.0, 0.0, 1.0], [45.0, 15.0, 0.0, 0.5]])    # N x 4: extra sources
    work = 64
    if case == "two_layer":
        cfg = decks.two_layer(3000)
    elif case == "two_layer_gates":
        cfg = decks.two_layer(3000, tend=2e-9, tstep=2e-10)
    elif case == "roulette":
        cfg = decks.two_layer(3000, minenergy=0.05)
    elif case == "savedet_all":
        cfg = decks.two_layer(6000, savedetflag=0x7F)
    elif case == "saveseed":
        cfg = decks.cube(6000, issaveseed=1)
    elif case == "energy":
        cfg = decks.two_layer(3000, outputtype="energy")
    elif case == "length":
        cfg = decks.two_layer(2000, outputtype="length")
    elif case == "multisrc_pick":
        cfg = decks.cube(4000, srcid=0, savedetflag=0x7F, **multi)
    elif case == "multisrc_volumes":
        cfg = decks.cube(4000, srcid=-1, **multi)
    elif case == "multisrc_fixed":
        cfg = decks.cube(3000, srcid=3, **multi)
    elif case == "gscatter":
        cfg = decks.two_layer(3000, gscatter=5)
    elif case == "saveref":
        cfg = decks.two_layer(3000, issaveref=1)
    elif case == "flat2d":
        cfg = decks.cube(3000)
        cfg.update(vol=np.ones((1, 60, 60), dtype=np.uint8), srcpos=[0.0, 29.0, 0.0], detpos=[[0, 19, 0, 2]])
    elif case == "skinvessel":
        cfg = benchmarks.get("skinvessel", 300)
    elif case == "phase_table":
        cfg = decks.cube(3000, invcdf=np.cos(np.linspace(np.pi, 0, 200)).astype(np.float32))
    elif case == "angle_table":
        cfg = decks.cube(3000, srcdir=[0, 0, 1, 0], angleinvcdf=np.linspace(0, 0.3, 16).astype(np.float32))
    elif case == "angle_table_discrete":
        cfg = decks.cube(3000, srcdir=[0, 0, 1, 1], angleinvcdf=np.array([0, 0.1, 0.2], dtype=np.float32))
    elif case == "rngdebug":
        cfg = decks.cube(100, debuglevel=1)
    else:
        cfg = decks.cube(3037)
        work = 100
    try:
        p, a, b = both(ref, port, cfg, work)
    except (KeyError, ValueError, TypeError) as e:
        pytest.skip("host config mirror does not take this option: %s" % e)
    assert_identical(a, b)
    if case.startswith("multisrc"):
        assert p.c.extrasrclen == 2
        if case == "multisrc_volumes":
            assert a["field"].size == 3 * 216000 and all(a["field"].reshape(3, -1).sum(axis=1) > 0)


# continuous media: oracle/_ref is the reference source built with -DMED_TYPE=99..104, the port decodes the words itself
@pytest.mark.parametrize("name", ["mua_float", "as_f2h", "label_half", "asgn_byte", "as_short"])
@pytest.mark.parametrize("reflect,det", [(0, 0), (1, 0), (1, 1)])
def test_port_bit_identical_continuous_media(ref, port, name, reflect, det):
    """updateproperty (src/mcx_core.cl:1079-1193) and the far-side index of the mismatch test (:3147) restated in C"""
    vol, prop, fmt = decks.media_volumes()[name]
    cfg = dict(benchmarks.get("cube60b", 3000), vol=vol, prop=prop, isreflect=reflect, issavedet=det, savedetflag="dxvw")
    p, a, b = both(ref, port, cfg)
    assert p.c.mediaformat == fmt and a["energytot"] == 3000
    assert_identical(a, b)


def test_port_bit_identical_continuous_media_with_a_cavity(ref, port):
    """zero words are background: a packet crossing an air gap inside a float-mua volume takes n from row 0 there (:1095)
    and is launched through the void in front of the tissue (skipvoid)"""
    mua = np.full((1, 60, 60, 60), 0.01, np.float32)
    mua[0, 20:40, 20:40, 25:32] = 0.0           # cavity
    mua[0, :, :, :4] = 0.0                      # air in front of the tissue: the pencil beam starts in it
    cfg = dict(benchmarks.get("cube60b", 3000), vol=mua, prop=decks.MEDIA_PROP2, isreflect=1, issavedet=0)
    p, a, b = both(ref, port, cfg)
    assert p.c.mediaformat == hostcfg.MEDIA_MUA_FLOAT and a["energytot"] == 3000
    assert_identical(a, b)


@pytest.mark.parametrize("case", ["cube60b", "no_reflection", "gates", "roulette", "hot_voxel"])
def test_port_bit_identical_rf_forward(ref, port, case):
    """complex packet weights of a forward run with a modulation frequency (src/mcx_core.cl:2427-2430, 2750-2760,
    2833-2841, 3035-3040), including the reference's quirk that the step whose real deposit spills past MAX_ACCUM loses
    its imaginary deposit (:2884-2893; `hot_voxel` crosses the limit several times)"""
    omega = 2 * np.pi * 100e6
    cfg = {
        "cube60b": dict(benchmarks.get("cube60b", 3000), omega=omega),
        "no_reflection": dict(benchmarks.get("cube60", 3000), omega=omega),
        "gates": dict(decks.two_layer(3000, tend=2e-9, tstep=2e-10), omega=omega),
        "roulette": dict(benchmarks.get("cube60b", 2000), omega=omega, minenergy=0.01, prop=[[0, 0, 1, 1], [0.05, 1.0, 0.01, 1.37]]),
        "hot_voxel": dict(benchmarks.get("cube60b", 6000), omega=omega, prop=[[0, 0, 1, 1], [0.0005, 1.0, 0.01, 1.37]]),
    }[case]
    p, a, b = both(ref, port, cfg)
    n = a["field"].size // 2
    assert p.rfplanes == 2 and a["field"].size == p.fieldlen
    assert a["field"][:n].sum() > 0 > a["field"][n:].sum()
    if case == "hot_voxel":
        assert a["field"][:n].max() > 2000.0
    assert_identical(a, b)


@pytest.mark.parametrize("case", ["rayleigh", "circular", "isotropic", "no_reflection", "two_layer", "rf_and_polarised"])
def test_port_bit_identical_polarised_light(ref, port, case):
    """Stokes vectors (src/mcx_core.cl:792-835): rejection sampling of the scattering angles from the Mueller matrix of the
    medium (:2454-2468), the vector rotated and renormalised at every event (:2562-2565), {I, Q, U, V} in the records
    (:917-922)"""
    two = decks.two_layer(3000)
    cfg = {
        "rayleigh": decks.pol_cfg(3000),
        "circular": decks.pol_cfg(3000, srciquv=[1, 0, 0, 1]),
        "isotropic": decks.pol_cfg(3000, smatrix=decks.isotropic_matrix(1)),
        "no_reflection": decks.pol_cfg(3000, isreflect=0),
        "two_layer": dict(two, smatrix=decks.rayleigh(len(two["prop"]) - 1), srciquv=[1, 0, 1, 0], issavedet=1, savedetflag="dpi",
                          detpos=[[30, 30, 0, 4]], maxdetphoton=3000),
        "rf_and_polarised": decks.pol_cfg(2000, omega=2 * np.pi * 100e6),
    }[case]
    p, a, b = both(ref, port, cfg)
    assert p.c.polmedianum >= 1 and a["detected"] > 0 and (a["detp"][:, -4] == 1.0).all()
    assert_identical(a, b)


@pytest.fixture(scope="module")
def replay_base(ref):
    import test_replay
    cfg = test_replay.baseline_cfg(60000)
    return cfg, ref.run(hostcfg.prepare(cfg), 1024, hostthreads=0)


@pytest.mark.parametrize("replaydet", [0, -1])
@pytest.mark.parametrize("ot", ["flux", "jacobian", "wltof", "wp", "wm", "wptof"])
def test_port_bit_identical_replay(ref, port, replay_base, ot, replaydet):
    """photon replay restated (src/mcx_core.cl:1590-1596 stream restart, :2568-2612 scattering-site outputs, :2845-2858
    absorption Jacobian), record indices exactly as the reference writes them -- including `f.w` instead of `f.w - 1` at
    scattering sites, which is why those outputs are handed tables padded by one element (tests/test_replay.py)"""
    import test_replay
    cfg, base = replay_base
    assert base["detected"] > 100
    p = hostcfg.prepare(test_replay.replay_cfg(cfg, base["detp"], base["seeds"], outputtype=ot, isnormalized=0, issaveseed=1, replaydet=replaydet))
    if ot in ("wp", "wm", "wptof"):
        test_replay.shift_records_for_the_reference(p)
    a = ref.run(p, 1, hostthreads=1)
    b = port.run(p, 1, hostthreads=1)
    assert a["field"].size == p.fieldlen and a["detected"] == p.c.nphoton
    assert a["energytot"] == b["energytot"] and a["energyesc"] == b["energyesc"]
    assert (bits(a["field"]) == bits(b["field"])).all() and np.abs(a["field"]).sum() > 0
    assert a["detected"] == b["detected"] and (bits(a["detp"]) == bits(b["detp"])).all() and (a["seeds"] == b["seeds"]).all()


@pytest.mark.parametrize("case", ["matched", "mismatched", "no_reflection", "tilted_surface", "detectors", "rf"])
def test_port_bit_identical_split_voxel_media(ref, port, case):
    """SVMC (MED_TYPE 97): updateproperty_svmc / ray_plane_intersect / reflectray_svmc (src/mcx_core.cl:1231-1344) and their
    call sites in the photon loop restated; the checker is the reference source built with -DMED_TYPE=97"""
    import test_gpu_svmc as sv
    surface = dict(vol=sv.tilted_slab(z0=30.4, sx=0.2, sy=0.1, below=1, above=0), prop=[[0, 0, 1, 1], [0.01, 1.0, 0.5, 1.37]])
    over = {
        "matched": dict(),
        "mismatched": dict(prop=[[0, 0, 1, 1], [0.02, 1.0, 0.8, 1.37], [0.005, 2.0, 0.9, 1.55]]),
        "no_reflection": dict(isreflect=0),
        "tilted_surface": surface,
        "detectors": dict(surface, issavedet=1, detpos=[[20, 20, 30.4, 4], [10, 20, 28.4, 3]], savedetflag="dsp", maxdetphoton=30000),
        "rf": dict(omega=2 * np.pi * 100e6),
    }[case]
    p, a, b = both(ref, port, sv.deck(4000, **over))
    assert p.c.mediaformat == 97 and a["energytot"] == 4000
    if case == "detectors":
        assert a["detected"] > 10
    assert_identical(a, b)


@pytest.mark.parametrize("case", ["M", "M_odd_split", "T_capped", "M_multisource"])
def test_port_bit_identical_trajectories(ref, port, case):
    """-D M / -D T: one record at the launch, at every scattering site and at the end of every packet (src/mcx_core.cl:
    929-948, 1497-1503, 2243-2249, 2625-2632), ids numbered exactly as the reference numbers them"""
    cfg = {
        "M": dict(benchmarks.get("cube60b", 3000), debuglevel="M", maxjumpdebug=500000),
        "M_odd_split": dict(benchmarks.get("cube60b", 3001), debuglevel="M", maxjumpdebug=500000),
        "T_capped": dict(benchmarks.get("cube60b", 3000), debuglevel="T", maxjumpdebug=1000),
        "M_multisource": dict(decks.cube(3000, srcpos=[[29.0, 29.0, 0.0, 1.0], [10.0, 40.0, 0.0, 1.0]], srcid=-1), debuglevel="M", maxjumpdebug=500000),
    }[case]
    p, a, b = both(ref, port, cfg)
    assert_identical(a, b)
    assert a["traj"].shape == b["traj"].shape and a["traj"].shape[0] == min(p.c.maxjumpdebug, a["traj"].shape[0]) >= 1000
    assert (bits(a["traj"]) == bits(b["traj"])).all()


@pytest.mark.parametrize("case", ["adjoint", "tilted", "srcid_minus_2", "adjoint_musp", "disk_over_the_edge"])
def test_port_bit_identical_adjoint_disk_sources(ref, port, case):
    """adjoint forward runs (and srcid == -2): the sources appended for the detectors start from a disk of the detector's
    radius (src/mcx_core.cl:1619, 2155-2183); the adjoint output types deposit like fluence (:2844)"""
    base = dict(benchmarks.get("cube60", 4000), issavedet=0, srcpos=[[30, 30, 1, 1], [30, 42, 1, 1]], srcdir=[[0, 0, 1, 0], [0, 0, 1, 0]],
                srcparam1=[[0, 0, 0, 0], [4, 0, 0, 0]], detpos=[[30, 42, 1, 4]], srcid=-1, outputtype="adjoint", isnormalized=0)
    cfg = {
        "adjoint": base,
        "tilted": dict(base, srcdir=[[0, 0, 1, 0], [0.3, -0.2, 0.93273790530888, 0]]),
        "srcid_minus_2": dict(base, srcid=-2, outputtype="fluence"),
        "adjoint_musp": dict(base, outputtype="adjoint_musp"),
        "disk_over_the_edge": dict(base, srcpos=[[30, 30, 1, 1], [2, 2, 1, 1]], detpos=[[2, 2, 1, 4]]),
    }[case]
    p, a, b = both(ref, port, cfg)
    vols = a["field"].reshape(2, 60, 60, 60)
    assert p.nsrcvol == 2 and vols[0].sum() > 0 and vols[1].sum() > 0
    if case == "adjoint":
        # the detector's volume enters through a disk, the source's through one voxel
        assert (vols[1, 0] > 0).sum() > 5 * (vols[0, 0] > 0.2 * vols[0, 0].max()).sum()
    assert_identical(a, b)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_port_bit_identical_on_random_combinations(ref, port, seed):
    """tools/oracle_fuzz.py: random combinations of source type, boundary codes, media, gates, detector flags and physics
    modes (RF forward, polarised light, trajectories, continuous media) -- 40 per seed, every output bit for bit"""
    import subprocess
    import sys
    root = os.path.dirname(HERE)
    out = subprocess.run([sys.executable, os.path.join(root, "tools", "oracle_fuzz.py"), "40", str(seed)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert " 0 differ" in out.stdout


def test_port_parallel_run_matches_serial_totals(port):
    p = hostcfg.prepare(benchmarks.get("cube60b", 2e4))
    a = port.run(p, 512, hostthreads=1)
    b = port.run(p, 512, hostthreads=0)
    assert a["energyesc"] == b["energyesc"] and a["detected"] == b["detected"]
    np.testing.assert_allclose(a["field"], b["field"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("deck,prefix", [("cube60", "17."), ("cube60b", "27.")])
def test_port_absorbed_fraction_pins(port, deck, prefix):
    """reference test/testmcx.sh:60-66"""
    o = port.run(hostcfg.prepare(benchmarks.get(deck, 1e5)), 1024, hostthreads=0)
    assert ("%.5f" % (100 * o["absorbed"])).startswith(prefix)
    assert abs(o["energytot"] - 1e5) <= 10
    if deck == "cube60b":
        assert o["detected"] == KAT["statistical"]["cube60b_detected_1e5_w1024"]
