"""CPU-side checks: the C ABI library loads and exports every symbol of include/mcxb200.h, the ctypes
mirror has the C layout, the seed table is glibc-rand() compatible, and the host configuration layer
(mcxcl_b200.hostcfg, mirror of mcx_initcfg/mcx_validatecfg/mcx_preprocess/mcx_maskdet) behaves like the
reference's.  No kernel is launched here."""
import ctypes as C
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

from mcxcl_b200 import abi, benchmarks, hostcfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcxb200.h")
ERR_ARG = -1            # MCXB_ERR_ARG


def test_library_exports_every_declared_symbol(lib):
    text = open(HEADER).read()
    declared = set(re.findall(r"\b(mcxb_[a-z0-9_]+)\s*\(", text))
    assert declared, "no prototypes found"
    bound = {name for name, _, _ in abi.SYMBOLS}
    assert declared == bound, (declared - bound, bound - declared)
    for name in declared:
        assert hasattr(lib, name)


def test_ctypes_layout_matches_c_header():
    src = r'''
#include <stdio.h>
#include <stddef.h>
#include "mcxb200.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu\n", sizeof(mcxb_config), sizeof(mcxb_output), sizeof(mcxb_gpuinfo), sizeof(mcxb_trace_step), sizeof(mcxb_source));
    printf("%zu %zu %zu %zu %zu %zu\n", offsetof(mcxb_config, vol), offsetof(mcxb_config, src), offsetof(mcxb_config, nphoton),
           offsetof(mcxb_config, bc), offsetof(mcxb_config, sched), offsetof(mcxb_config, accum));
    printf("%zu %zu %zu\n", offsetof(mcxb_output, energytot), offsetof(mcxb_output, kernel_launches), offsetof(mcxb_output, stats));
    return 0;
}'''
    with tempfile.TemporaryDirectory() as d:
        cfile = os.path.join(d, "t.c")
        open(cfile, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), cfile, "-o", exe])
        lines = subprocess.check_output([exe]).decode().split("\n")
    sizes = [int(x) for x in lines[0].split()]
    assert sizes == [C.sizeof(abi.Config), C.sizeof(abi.Output), C.sizeof(abi.GPUInfo), C.sizeof(abi.TraceStep), C.sizeof(abi.Source)]
    offs = [int(x) for x in lines[1].split()]
    assert offs == [getattr(abi.Config, f).offset for f in ("vol", "src", "nphoton", "bc", "sched", "accum")]
    offs = [int(x) for x in lines[2].split()]
    assert offs == [getattr(abi.Output, f).offset for f in ("energytot", "kernel_launches", "stats")]


def test_ctypes_mirror_matches_every_field_offset():
    """every member of every struct the Python mirror declares sits where the C header puts it (offset and size), so a
    field appended on one side only cannot go unnoticed"""
    pairs = [("mcxb_config", abi.Config), ("mcxb_output", abi.Output), ("mcxb_gpuinfo", abi.GPUInfo), ("mcxb_trace_step", abi.TraceStep),
             ("mcxb_source", abi.Source)]
    body = []
    for cname, cls in pairs:
        for fname, _ in cls._fields_:
            body.append('printf("%%zu %%zu\\n", offsetof(%s, %s), sizeof(((%s*)0)->%s));' % (cname, fname, cname, fname))
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "mcxb200.h"\nint main(void) {\n' + "\n".join(body) + "\nreturn 0;\n}\n"
    with tempfile.TemporaryDirectory() as d:
        cfile = os.path.join(d, "t.c")
        open(cfile, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), cfile, "-o", exe])
        got = [tuple(int(x) for x in ln.split()) for ln in subprocess.check_output([exe]).decode().split("\n") if ln]
    want = [(getattr(cls, fname).offset, getattr(cls, fname).size) for _, cls in pairs for fname, _ in cls._fields_]
    names = ["%s.%s" % (cname, fname) for cname, cls in pairs for fname, _ in cls._fields_]
    assert len(got) == len(want)
    assert [n for n, g, w in zip(names, got, want) if g != w] == []


def _cli_dump(name):
    """the reference's own CLI, relinked (integration/_build/mcxcl): `--dumpjson -` prints the configuration AFTER
    mcx_validatecfg / mcx_preprocess (src/mcx_utils.c:1447-1712), without touching a device"""
    import json
    exe = os.path.join(ROOT, "integration", "_build", "mcxcl")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/mcxcl not built (python integration/build_cli.py needs /root/reference)")
    decks = {"digimouse": "example/digimouse/digimouse.json", "qtest": "example/quicktest/qtest.inp"}      # not built in: the shipped decks
    if name in decks:
        path = os.path.join("/root/reference", decks[name])
        if not os.path.exists(path):
            pytest.skip("the reference tree is not here")
        args, cwd = ["-f", os.path.basename(path)], os.path.dirname(path)
    else:
        args, cwd = ["--bench", name], None
    out = subprocess.run([exe] + args + ["--dumpjson", "-", "-n", "1000"], capture_output=True, text=True, timeout=300, cwd=cwd)
    assert out.returncode == 0, out.stderr[-500:]
    return json.loads(out.stdout)


@pytest.mark.parametrize("name", ["cube60", "cube60b", "cube60planar", "skinvessel", "colin27", "digimouse", "qtest"])
def test_host_mirror_matches_the_reference_clis_own_preprocessing(name):
    """mcxcl_b200.hostcfg + mcxcl_b200.benchmarks (what the GPU tests and bench.py feed the engine) against the reference's
    code itself: built-in benchmark -> mcx_validatecfg -> mcx_preprocess -> JSON dump.  Source records (launch voxel and
    label bit patterns in Param2 included), media table, gates, detectors and flags agree bit for bit; colin27's volume
    voxel for voxel."""
    import base64
    import zlib
    d = _cli_dump(name)
    cfg = benchmarks.get(name, 1000)
    p = hostcfg.prepare(cfg)
    c = p.c
    bits = lambda x: np.asarray(x, np.float32).view(np.uint32)          # noqa: E731
    src, ses = d["Optode"]["Source"], d["Session"]
    assert ses["Photons"] == c.nphoton
    if name != "digimouse":            # the shipped digimouse deck asks for seed 2147483647; the bench uses one seed for every deck
        assert ses["RNGSeed"] == c.seed
    assert hostcfg.SRCTYPES.index(src["Type"]) == c.srctype
    assert np.array_equal(bits(src["Pos"]), bits([c.src.pos.x, c.src.pos.y, c.src.pos.z]))
    assert np.array_equal(bits(src["Dir"]), bits([c.src.dir.x, c.src.dir.y, c.src.dir.z, c.src.dir.w]))
    assert np.array_equal(bits(src["Param1"]), bits([c.src.param1.x, c.src.param1.y, c.src.param1.z, c.src.param1.w]))
    assert np.array_equal(bits(src["Param2"]), bits([c.src.param2.x, c.src.param2.y, c.src.param2.z, c.src.param2.w]))
    assert list(d["Domain"]["Dim"]) == list(p.dims)
    assert np.array_equal(bits([d["Forward"]["T0"], d["Forward"]["T1"], d["Forward"]["Dt"]]), bits([c.tstart, c.tend, c.tstep]))
    assert np.float32(d["Domain"]["LengthUnit"]) == np.float32(c.unitinmm)
    # the dump writes mua / mus per mm again (divided by the voxel size); the table at the boundary is per voxel edge
    med = np.array([[m["mua"], m["mus"], m["g"], m["n"]] for m in d["Domain"]["Media"]], np.float64)
    mine = np.array([[c.prop[i].x, c.prop[i].y, c.prop[i].z, c.prop[i].w] for i in range(c.medianum)], np.float64)
    assert med.shape == mine.shape
    med[:, :2] *= float(np.float32(c.unitinmm))
    if c.unitinmm == 1.0:
        assert np.array_equal(bits(med), bits(mine))
    else:
        np.testing.assert_allclose(mine, med, rtol=3e-7)
    det = d["Optode"].get("Detector", [])
    assert len(det) == c.detnum
    for i, row in enumerate(det):
        assert np.array_equal(bits(list(row["Pos"]) + [row["R"]]), bits([c.detpos[i].x, c.detpos[i].y, c.detpos[i].z, c.detpos[i].w]))
    assert bool(ses["DoMismatch"]) == bool(c.isreflect) and bool(ses["DoPartialPath"]) == bool(c.issavedet)
    assert bool(ses["DoSpecular"]) == (c.isspecular > 0) and bool(ses["DoNormalize"]) == bool(c.isnormalized)
    if c.issavedet:
        assert ses["SaveDataMask"] == c.savedetflag
    sh = d.get("Shapes")
    if isinstance(sh, dict) and "_ArrayZipData_" in sh:                 # volume benchmarks carry the volume as a JData array
        raw = np.frombuffer(zlib.decompress(base64.b64decode(sh["_ArrayZipData_"])), dtype=np.dtype(sh["_ArrayType_"]))
        assert np.array_equal(raw.reshape(sh["_ArraySize_"]), np.asarray(cfg["vol"]))


def _json_number(x):
    """JData spelling of the non-finite floats the reference's JSON reader understands"""
    x = float(x)
    if np.isnan(x):
        return "_NaN_"
    if np.isinf(x):
        return "_Inf_" if x > 0 else "-_Inf_"
    return x


def _from_json_number(x):
    if isinstance(x, str):
        return float(x.replace("-_Inf_", "-inf").replace("_Inf_", "inf").replace("_NaN_", "nan"))
    return x


@pytest.mark.parametrize("name", sorted(__import__("decks").SOURCES))
def test_source_records_match_the_reference_clis_own_preprocessing(name, tmp_path):
    """every source deck of tests/decks.py written as a JSON input file, read by the reference's parser (mcx_loadjson) and
    preprocessed by its own code (src/mcx_utils.c:1447-1712: direction normalisation, 0-based positions, launch voxel and
    label, the per-type parameter conventions), against what hostcfg.prepare() hands the engine: bit for bit"""
    import decks
    import json
    exe = os.path.join(ROOT, "integration", "_build", "mcxcl")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/mcxcl not built (python integration/build_cli.py needs /root/reference)")
    cfg = decks.cube(1000, **decks.SOURCES[name])
    pad = lambda v: [float(x) for x in v] + [0.0] * (4 - len(v))          # noqa: E731
    src = {"Type": cfg.get("srctype", "pencil"), "Pos": [float(x) for x in cfg["srcpos"]], "Dir": [_json_number(x) for x in cfg.get("srcdir", [0, 0, 1])]}
    for key, field in (("srcparam1", "Param1"), ("srcparam2", "Param2")):
        if key in cfg:
            src[field] = pad(cfg[key])
    if "srcpattern" in cfg:
        pat = np.asarray(cfg["srcpattern"], np.float32)
        src["Pattern"] = dict(zip(("Nx", "Ny", "Nz"), pat.shape), Data=[float(x) for x in pat.ravel(order="F")])
    deck = {"Session": {"ID": "deck", "Photons": 1000, "RNGSeed": int(cfg["seed"]), "DoMismatch": bool(cfg["isreflect"]), "DoSpecular": bool(cfg.get("isspecular", 0))},
            "Forward": {"T0": cfg["tstart"], "T1": cfg["tend"], "Dt": cfg["tstep"]},
            "Domain": {"OriginType": 1, "LengthUnit": 1, "Dim": [60, 60, 60],
                       "Media": [dict(zip(("mua", "mus", "g", "n"), [float(x) for x in row])) for row in cfg["prop"]]},
            "Optode": {"Source": src, "Detector": [{"Pos": [float(x) for x in d[:3]], "R": float(d[3])} for d in cfg["detpos"]]},
            "Shapes": [{"Grid": {"Tag": 1, "Size": [60, 60, 60]}}]}
    with open(os.path.join(tmp_path, "deck.json"), "w") as f:
        json.dump(deck, f)
    out = subprocess.run([exe, "-f", "deck.json", "--dumpjson", "-"], capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert out.returncode == 0, out.stderr[-500:]
    got = json.loads(out.stdout)["Optode"]["Source"]
    c = hostcfg.prepare(cfg).c
    bits = lambda x: np.asarray([_from_json_number(v) for v in x], np.float32).view(np.uint32)          # noqa: E731
    assert hostcfg.SRCTYPES.index(got["Type"]) == c.srctype
    assert np.array_equal(bits(got["Pos"]), bits([c.src.pos.x, c.src.pos.y, c.src.pos.z]))
    mydir = [c.src.dir.x, c.src.dir.y, c.src.dir.z, c.src.dir.w]
    if got["Dir"][3] is None:           # the dump has no spelling for NaN / -inf (isotropic / Lambertian launch): compare the vector part
        assert not np.isfinite(mydir[3]) and np.array_equal(bits(got["Dir"][:3]), bits(mydir[:3]))
    else:
        assert np.array_equal(bits(got["Dir"]), bits(mydir))
    assert np.array_equal(bits(got["Param1"]), bits([c.src.param1.x, c.src.param1.y, c.src.param1.z, c.src.param1.w]))
    assert np.array_equal(bits(got["Param2"]), bits([c.src.param2.x, c.src.param2.y, c.src.param2.z, c.src.param2.w]))


@pytest.mark.parametrize("name", ["cube60b", "colin27", "qtest"])
def test_detector_mask_matches_the_reference_clis_own(name, tmp_path):
    """mcx_maskdet (src/mcx_utils.c:4050-4190) marks the tissue voxels on the surface under every detector with the top bit;
    `-M 1` makes the reference's CLI write the masked volume and exit.  hostcfg.maskdet marks the same voxels."""
    import base64
    import json
    import zlib
    exe = os.path.join(ROOT, "integration", "_build", "mcxcl")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/mcxcl not built (python integration/build_cli.py needs /root/reference)")
    if name == "qtest":
        src = "/root/reference/example/quicktest"
        if not os.path.exists(src):
            pytest.skip("the reference tree is not here")
        for f in ("qtest.inp", "cubic60.json"):
            open(os.path.join(tmp_path, f), "w").write(open(os.path.join(src, f)).read())
        args = ["-f", "qtest.inp", "-s", "qtest"]
    else:
        args = ["--bench", name]
    out = subprocess.run([exe] + args + ["-n", "1000", "-M", "1"], capture_output=True, text=True, cwd=tmp_path, timeout=300)
    assert out.returncode == 0 and "volume mask is saved" in out.stdout, (out.stdout[-300:], out.stderr[-300:])
    files = [f for f in os.listdir(tmp_path) if f.endswith("_vol.jnii")]
    assert len(files) == 1
    nd = json.load(open(os.path.join(tmp_path, files[0])))["NIFTIData"]
    theirs = np.frombuffer(zlib.decompress(base64.b64decode(nd["_ArrayZipData_"])), dtype=np.dtype(nd["_ArrayType_"])).reshape(nd["_ArraySize_"])
    p = hostcfg.prepare(benchmarks.get(name, 1000))
    nx, ny, nz = p.dims
    mine = np.asarray(p.keep["vol"], np.uint32).reshape(nz, ny, nx).transpose(2, 1, 0)         # x fastest at the boundary
    assert theirs.shape == mine.shape
    assert (theirs >> 31).sum() > 8 and np.array_equal(theirs, mine)


def test_flag_parsers_match_the_reference_clis_own():
    """-w (saveflag[], src/mcx_utils.c:134), -D (debugflag[]) and -O (outputtype[]) as the reference's CLI reads them, against
    the host mirror's parsers"""
    import json
    exe = os.path.join(ROOT, "integration", "_build", "mcxcl")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/mcxcl not built (python integration/build_cli.py needs /root/reference)")

    def session(args):
        out = subprocess.run([exe, "--bench", "cube60b", "-n", "1000", "--dumpjson", "-"] + args, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr[-300:]
        return json.loads(out.stdout)["Session"]

    base = benchmarks.get("cube60b", 1000)
    for w in ("DP", "DSPMXVW", "dxv", "W", "SP", "DPXVW", "M"):
        assert session(["-w", w])["SaveDataMask"] == hostcfg.prepare(dict(base, savedetflag=w)).c.savedetflag
    for dbg in ("R", "M", "P", "T", "MP"):
        assert session(["-D", dbg])["DebugFlag"] == hostcfg.prepare(dict(base, debuglevel=dbg, maxjumpdebug=1000)).c.debuglevel
    for ot in "XFEL":
        assert hostcfg.OUTPUTTYPES[session(["-O", ot])["OutputType"]] == hostcfg.prepare(dict(base, outputtype=ot.lower())).c.outputtype


def test_seed_table_is_glibc_rand(lib):
    """src/mcx_host.cpp:696-700, 759-768: srand(seed); seeds[i] = rand().  Compared with the C library itself."""
    libc = C.CDLL("libc.so.6")
    libc.rand.restype = C.c_int
    for seed, skip, n in ((1648335518, 0, 64), (29012392, 17, 33), (1, 0, 8), (2147483647, 3, 5)):
        libc.srand(C.c_uint(seed))
        want = np.array([libc.rand() for _ in range(4 * (skip + n))], dtype=np.uint32)[4 * skip:]
        got = np.zeros(4 * n, dtype=np.uint32)
        lib.mcxb_fill_seeds(seed, skip, n, got.ctypes.data)
        assert (got == want).all(), seed


def test_seed_slices_concatenate(lib):
    """rank r of a multi-GPU job skips r*nthread records of ONE stream (src/mcx_host.cpp:759-768)."""
    whole = np.zeros(4 * 96, dtype=np.uint32)
    lib.mcxb_fill_seeds(12345, 0, 96, whole.ctypes.data)
    parts = []
    for r in range(3):
        part = np.zeros(4 * 32, dtype=np.uint32)
        lib.mcxb_fill_seeds(12345, 32 * r, 32, part.ctypes.data)
        parts.append(part)
    assert (np.concatenate(parts) == whole).all()


def test_prepare_cube60_matches_reference_preprocessing():
    p = hostcfg.prepare(benchmarks.get("cube60b", 1e6))
    c = p.c
    assert (c.dimx, c.dimy, c.dimz, c.medianum) == (60, 60, 60, 3)
    assert p.maxgate == 1 and p.fieldlen == 216000
    assert c.savedetflag == 5 and p.partialdata == 2 and p.reclen == 3
    # launch voxel index / label stored as uint bits in param2.z/.w (src/mcx_utils.c:1718-1749)
    z, w = np.array([c.src.param2.z, c.src.param2.w], dtype=np.float32).view(np.uint32)
    assert (z, w) == (29 * 60 + 29, 1)
    vol = p.keep["vol"]
    assert int((vol >> 31).sum()) == 48          # mcxcl --bench cube60 --dumpmask flags 48 voxels
    assert set(np.unique(vol & 0x7FFFFFFF)) == {1}
    assert c.isreflect == 1 and hostcfg.prepare(benchmarks.get("cube60", 1e3)).c.isreflect == 0


def test_prepare_unitinmm_and_zero_mus():
    p = hostcfg.prepare(benchmarks.get("skinvessel", 1e3))
    prop = p.keep["prop"]
    assert prop[0, 1] == np.float32(1e-10)                       # mus==0 -> EPS (src/mcx_utils.c:1647-1653)
    np.testing.assert_allclose(prop[2, 0], np.float32(23.05426549) * np.float32(0.005), rtol=1e-7)
    assert p.c.issavedet == 0 and p.c.savedetflag == 0 and p.c.srctype == 8
    assert sorted(np.unique(p.keep["vol"]).tolist()) == [1, 2, 3, 4]
    counts = np.bincount(p.keep["vol"])
    assert counts[1] == 800000 and counts[4] == 480000 and counts[2] == 251400


def test_prepare_one_based_coordinates():
    p = hostcfg.prepare(benchmarks.get("qtest", 1e3))
    assert (p.c.src.pos.x, p.c.src.pos.y, p.c.src.pos.z) == (29.0, 29.0, 0.0)
    np.testing.assert_array_equal(p.keep["detpos"][0], np.array([29, 19, 0, 1], dtype=np.float32))


def test_prepare_direction_normalised():
    cfg = benchmarks.get("cube60", 10)
    cfg["srcdir"] = [0.1636, 0.4569, -0.8743]
    p = hostcfg.prepare(cfg)
    d = np.array([p.c.src.dir.x, p.c.src.dir.y, p.c.src.dir.z], dtype=np.float64)
    assert abs(np.linalg.norm(d) - 1) < 1e-6


@pytest.mark.parametrize("mutate,code", [
    (lambda c: c.pop("prop"), -4),
    (lambda c: c.update(vol=np.ones((4, 4), np.uint8)), -4),
    (lambda c: c.update(srcdir=[0, 0, 0]), -4),
    (lambda c: c.update(prop=[[0, 0, 1, 1]]), -4),
    (lambda c: c.update(tend=0.0), -6),
    (lambda c: c.update(srctype="laser"), -6),
    (lambda c: c.update(bc="xyzabc"), -4),
    (lambda c: c.update(respin=0), -1),              # src/mcx_utils.c:1628-1630
])
def test_prepare_rejects_like_reference(mutate, code):
    cfg = benchmarks.get("cube60", 10)
    mutate(cfg)
    with pytest.raises(hostcfg.ConfigError) as e:
        hostcfg.prepare(cfg)
    assert e.value.code == code and "MCXCL ERROR(%d)" % code in str(e.value)


def test_boundary_condition_parsing():
    codes = hostcfg.parse_bc("aarraa")
    assert codes[:6].tolist() == [2, 2, 1, 1, 2, 2] and not codes[6:].any()
    codes = hostcfg.parse_bc("______111111")
    assert codes[:6].tolist() == [0] * 6 and codes[6:].tolist() == [1] * 6
    assert hostcfg.parse_savedetflag("dspxvw") == 0x77 and hostcfg.parse_savedetflag("DP") == 5


def test_multisource_rows():
    cfg = benchmarks.get("cube60", 10)
    cfg.update(srcpos=[[29, 29, 0], [10, 10, 0], [50, 50, 0]], srcdir=[[0, 0, 1]] * 3, srcid=-1)
    p = hostcfg.prepare(cfg)
    assert p.c.extrasrclen == 2 and p.nsrcvol == 3 and p.fieldlen == 3 * 216000
    assert p.keep["extra"].shape == (2, 16)
    z = p.keep["extra"][:, 14].view(np.uint32)
    assert z.tolist() == [10 * 60 + 10, 50 * 60 + 50]


def test_engine_refuses_to_run_without_a_gpu(lib):
    """no CPU fallback: on a box without CUDA devices the product path fails loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mcxcl_b200 import engine
    with pytest.raises(RuntimeError) as e:
        engine.run(benchmarks.get("cube60", 100))
    assert "failed" in str(e.value)
    assert engine.gpuinfo() == []


def _create_error(lib, cfg):
    import ctypes as C
    p = hostcfg.prepare(cfg) if isinstance(cfg, dict) else cfg
    sim = C.c_void_p()
    rc = lib.mcxb_sim_create(C.byref(p.c), 0, C.byref(sim))
    assert not sim.value
    return rc, (lib.mcxb_last_error() or b"").decode()


def test_c_abi_argument_errors_are_reported_before_any_device_work(lib):
    """the refusals of include/mcxb200.h are argument checks: code MCXB_ERR_ARG (-1) and a message that names the reason,
    with or without a GPU (the messages follow the reference's where it has one: src/mcx_utils.c:1633-1636, 1660,
    src/mcx_host.cpp:723)"""
    import ctypes as C
    base = benchmarks.get("cube60", 100)
    # the host mirror (like mcx_validatecfg) refuses most of these first: set the struct members behind its back
    cases = [
        (dict(respin=-2), "respin"),
        (dict(tstep=0.0), "time gate"),
        (dict(outputtype=3), "replay"),              # MCXB_OT_JACOBIAN
        (dict(debuglevel=abi.DEBUG_MOVE, maxjumpdebug=0), "maxjumpdebug"),
        (dict(srctype=99), "source type"),
        (dict(srcnum=2), "photon sharing"),
        (dict(extrasrclen=1), "srcdata"),
        (dict(dimx=40000), "grid dimensions"),
    ]
    for fields, word in cases:
        p = hostcfg.prepare(base)
        for k, v in fields.items():
            setattr(p.c, k, v)
        rc, msg = _create_error(lib, p)
        assert rc == ERR_ARG and word in msg, (fields, rc, msg)
    # optical properties the kernel cannot leave again: refused instead of hanging the device
    for row in ([0.005, float("nan"), 0.01, 1.37], [0.005, 1.0, 0.01, 0.0], [float("inf"), 1.0, 0.01, 1.37]):
        p = hostcfg.prepare(base)
        p.keep["badprop"] = np.array([[0, 0, 1, 1], row], np.float32)
        p.c.prop = p.keep["badprop"].ctypes.data_as(type(p.c.prop))
        rc, msg = _create_error(lib, p)
        assert rc == ERR_ARG and "media row 1" in msg, (row, rc, msg)
    p = hostcfg.prepare(base)
    p.c.src.pos.y = float("nan")
    rc, msg = _create_error(lib, p)
    assert rc == ERR_ARG and "source 0" in msg
    p = hostcfg.prepare(base)
    p.c.abi_version = abi.ABI_VERSION - 1
    rc, msg = _create_error(lib, p)
    assert rc == ERR_ARG and "abi_version" in msg
    # the multi-device call: device list checks
    p = hostcfg.prepare(base)
    out = abi.Output()
    for devs, word in (([], "devices"), ([0, 1, 0], "listed twice"), (list(range(abi.MAX_DEVICES + 1)), "devices")):
        arr = (C.c_int * max(1, len(devs)))(*devs)
        rc = lib.mcxb_run_simulation_multi(C.byref(p.c), arr, len(devs), None, C.byref(out), None)
        assert rc == ERR_ARG and word in (lib.mcxb_last_error() or b"").decode()
    # adjoint products: volumes of at least one source and one detector
    buf = np.zeros(8, np.float32)
    rc = lib.mcxb_adjoint_products(0, buf.ctypes.data, None, 2, 2, 2, 1, 0, 1, 0, buf.ctypes.data)
    assert rc == ERR_ARG and "at least one source" in (lib.mcxb_last_error() or b"").decode()


def test_photon_split_of_the_multi_gpu_call_matches_the_host_mirror(lib):
    """mcxb_split_photons (C ABI, used by mcxb_run_simulation_multi) == multigpu.split_photons (one process per GPU):
    the reference's workload ratio with the remainder given to the first devices (src/mcx_host.cpp:1011-1012)"""
    import ctypes as C
    from mcxcl_b200 import multigpu
    for nph, wl in ((10, [1, 1, 1]), (1000000000, [1] * 8), (1000, [3, 1]), (7, [0, 1, 1]), (400001, [3, 1]), (0, [1, 1])):
        share = (C.c_uint64 * len(wl))()
        lib.mcxb_split_photons(nph, (C.c_float * len(wl))(*wl), len(wl), share)
        assert list(share) == multigpu.split_photons(nph, wl)
    share = (C.c_uint64 * 3)()
    lib.mcxb_split_photons(10, None, 3, share)                  # no weights: equal shares
    assert list(share) == [4, 3, 3]


def test_nccl_is_bound_at_run_time(lib):
    """no link-time dependency on NCCL: the library loads without it and reports the version it finds by dlopen"""
    import subprocess
    out = subprocess.run(["ldd", os.path.join(ROOT, "mcxcl_b200", "libmcxb200.so")], capture_output=True, text=True).stdout
    assert "nccl" not in out
    assert lib.mcxb_nccl_version() >= 20000


def test_continuous_media_packing_follows_pmcxcl():
    """hostcfg.pack_continuous_volume mirrors the packing of src/pmcxcl.cpp:108-400 (what reaches the boundary in
    Config.vol / Config.mediabyte)"""
    mua = np.full((3, 4, 5), 0.005, np.float32)
    mus = np.full((3, 4, 5), 1.5, np.float32)
    w, fmt = hostcfg.pack_continuous_volume(mua[None], 2.0)
    assert fmt == hostcfg.MEDIA_MUA_FLOAT and w.dtype == np.uint32 and (w.view(np.float32) == np.float32(0.01)).all()
    z = mua.copy()
    z[0, 0, 0] = 0.0
    z[1, 1, 1] = np.nan
    w, _ = hostcfg.pack_continuous_volume(z[None], 1.0)
    assert w[0, 0, 0] == np.float32(1.19209290e-07).view(np.uint32) and w[1, 1, 1] == 0         # zero mua is not a 0-label voxel; NaN is
    w, fmt = hostcfg.pack_continuous_volume(np.stack([mua, mus]), 1.0)
    assert fmt == hostcfg.MEDIA_AS_F2H
    lo = (w & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float32)
    hi = (w >> 16).astype(np.uint16).view(np.float16).astype(np.float32)
    assert (hi == 1.5).all() and (np.abs(lo - 0.005) < 0.005 * 2.0 ** -10).all() and (lo <= 0.005).all()   # mantissa truncated, not rounded
    # exactly representable halves survive unchanged; the conversion never rounds up
    vals = np.array([1.0, 0.5, 0.375, 2.0, 1000.0, 6.1035156e-05, 65504.0], np.float32)
    assert (hostcfg._float_to_half_bits(vals).astype(np.uint16).view(np.float16).astype(np.float32) == vals).all()
    rs = np.random.RandomState(5)
    f = rs.uniform(1e-3, 50, 2000).astype(np.float32)
    h = hostcfg._float_to_half_bits(f).astype(np.uint16).view(np.float16).astype(np.float32)
    assert (h <= f).all() and (f - h < f * 2.0 ** -10).all()
    b = (np.arange(4 * 3 * 4 * 5) % 251).astype(np.uint8).reshape(4, 3, 4, 5)
    w, fmt = hostcfg.pack_continuous_volume(b)
    assert fmt == hostcfg.MEDIA_ASGN_BYTE and (w.view(np.uint8).reshape(3, 4, 5, 4) == np.moveaxis(b, 0, -1)).all()
    s16 = (np.arange(2 * 3 * 4 * 5) * 1021 % 65521).astype(np.uint16).reshape(2, 3, 4, 5)
    w, fmt = hostcfg.pack_continuous_volume(s16)
    assert fmt == hostcfg.MEDIA_AS_SHORT and ((w & 0xFFFF) == s16[0]).all() and ((w >> 16) == s16[1]).all()
    lh = np.zeros((3, 3, 4, 5), np.float32)
    lh[0], lh[1], lh[2] = 0.5, 1, 7                       # value 0.5 replaces slot 1 (mus, scaled by unitinmm) of label 7
    w, fmt = hostcfg.pack_continuous_volume(lh, 2.0)
    assert fmt == hostcfg.MEDIA_LABEL_HALF and ((w & 0x3FFF) == 7).all() and (((w >> 14) & 3) == 1).all()
    assert ((w >> 16).astype(np.uint16).view(np.float16) == np.float16(1.0)).all()
    p = hostcfg.prepare(dict(benchmarks.get("cube60", 10), vol=np.full((1, 60, 60, 60), 0.005, np.float32), issavedet=0))
    assert p.c.mediaformat == 101 and p.dims == (60, 60, 60)
    with pytest.raises(hostcfg.ConfigError):
        hostcfg.prepare(dict(benchmarks.get("cube60", 10), vol=np.zeros((4, 6, 6, 6), np.uint8), prop=[[0, 0, 1, 1], [0.005, 1, 0, 1.37]]))
