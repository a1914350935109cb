"""The reference's UNCHANGED command-line front-end (src/mcxcl.c + src/mcx_utils.c ...), relinked against the B200
engine through integration/mcx_cuda_host.cpp, run with the commands and known answers of the reference's own
test script (test/testmcx.sh:60-132).  Output files written by the reference's unchanged writers are read back
and compared with the Python host mirror."""
import json
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from mcxcl_b200 import benchmarks, engine, hostcfg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MCX = os.path.join(ROOT, "integration", "_build", "mcxcl")
ANSI = re.compile(r"\x1b\[[0-9;]*m")


def mcx(args, cwd):
    if not os.path.exists(MCX):
        pytest.skip("integration/_build/mcxcl not built (python integration/build_cli.py needs /root/reference)")
    r = subprocess.run([MCX] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    return r.returncode, ANSI.sub("", r.stdout + r.stderr)


def absorbed(out):
    m = re.search(r"absorbed:\s*([0-9.]+)%", out)
    assert m, out[-2000:]
    return float(m.group(1))


def test_gpu_listing(tmp_path):
    """test/testmcx.sh:32-34: `mcxcl -L | grep 'Global [Mm]emory'`"""
    rc, out = mcx(["-L"], tmp_path)
    assert rc == 0 and re.search(r"Global [Mm]emory", out) and "B200" in out


# (arguments, regex the reference's script greps for) -- test/testmcx.sh:60-114
PINS = [
    (["--bench", "cube60", "-S", "0"], r"absorbed:.*17\.[0-9]+%"),
    (["--bench", "cube60b", "-S", "0"], r"absorbed:.*27\.[0-9]+%"),
    (["--bench", "cube60", "-b", "1", "-S", "0"], r"absorbed:.*27\.[0-9]+%"),
    (["--bench", "cube60", "-b", "0", "-B", "aarraa", "-S", "0"], r"absorbed:.*27\.[0-9]+%"),
    (["--bench", "cube60", "--bc", "cccccc", "-n", "1e3", "-d", "0", "-S", "0"], r"absorbed:.*99\.[0-9]+%"),
    (["--bench", "cube60b"], r"detected.*4[0-9]+ photons"),
    (["--bench", "cube60planar"], r"absorbed:.*25\.[0-9]+%"),
    (["--bench", "cube60", "--json", '{"Optode":{"Source":{"Type":"isotropic","Pos":[29,29,29]}}}', "-d", "0", "-S", "0"], r"absorbed:.*88\.[0-9]+%"),
    (["--bench", "cube60", "--json", '{"Domain":{"Media":[[0,0,1,1],[0.001,0.001,0,1]]},"Optode":{"Source":{"Type":"cone","Param1":[0.5,0,0,0]}}}', "-d", "0", "-S", "0"], r"absorbed:.*6\.[0-9]+%"),
    (["--bench", "cube60planar", "--json", '{"Optode":{"Source":{"Type":"fourier","Param1":[40,0,0,2]}}}', "-d", "0", "-S", "0"], r"absorbed:.*25\.[0-9]+%"),
    (["--bench", "cube60planar", "--json", '{"Optode":{"Source":{"Type":"pencilarray","Param1":[40,0,0,4],"Param2":[0,20,0,2]}}}', "-d", "0", "-S", "0"], r"absorbed:.*23\.[0-9]+%"),
    (["--bench", "cube60b", "--json", '{"Shapes":[{"Grid":{"Tag":1,"Size":[1,100,100]}},{"Box":{"Tag":2,"O":[0,30,10],"Size":[1,40,40]}}],"Domain":{"Media":[[0,0,1,1],[0.02,0.1,0.9,1.37],[0.02,10,0.9,6.85]]},"Optode":{"Source":{"Pos":[0,50,0],"Dir":[0,0,1]}}}', "-d", "0", "-S", "0"], r"absorbed:.*6[0-9]\.[0-9]+%"),
    (["--bench", "skinvessel", "-S", "0"], r"absorbed:.*39\.[0-9]+%"),
    (["--bench", "spherebox", "-S", "0"], r"absorbed:.*1[01]\.[0-9]+%"),
]


@pytest.mark.parametrize("k", range(len(PINS)))
def test_reference_script_pins(tmp_path, k):
    args, pattern = PINS[k]
    rc, out = mcx(args + ([] if "-n" in args else ["-n", "1e5"]), tmp_path)
    assert rc == 0, out[-2000:]
    assert re.search(pattern, out), out[-1500:]


def test_boundary_detector_flags(tmp_path):
    """test/testmcx.sh:104-106: --bc ______111111 -n 1e4 detects 97xx-99xx photons"""
    rc, out = mcx(["--bench", "cube60", "--bc", "______111111", "-n", "1e4"], tmp_path)
    m = re.search(r"detected\s+([0-9]+) photons", out)
    assert rc == 0 and m and 9700 <= int(m.group(1)) <= 9999


def test_saving_photon_seeds_jnii(tmp_path):
    """test/testmcx.sh:116-118: `--bench cube60 -q 1 -F jnii -S 0` prints 'after encoding: 13x.x%' for the seed block"""
    rc, out = mcx(["--bench", "cube60", "-q", "1", "-F", "jnii", "-S", "0", "-n", "1e5"], tmp_path)
    assert rc == 0 and re.search(r"after encoding: 13[0-9]\.[0-9]+%", out), out[-1500:]


def test_photon_replay_through_the_cli(tmp_path):
    """test/testmcx.sh:120-128: a baseline run saves seeds into replaytest_detp.jdat, `-E` replays them: the replay
    simulates and detects exactly the photons the baseline detected, and absorbs 3[0-8].x %"""
    rc, out1 = mcx(["--bench", "cube60", "-s", "replaytest", "-q", "1", "-S", "0", "-n", "1e5"], tmp_path)
    assert rc == 0, out1[-2000:]
    rc, out2 = mcx(["--bench", "cube60", "-E", "replaytest_detp.jdat", "-S", "0", "-n", "1e5"], tmp_path)
    assert rc == 0, out2[-2000:]
    nums = re.findall(r"(?:simulated|detected)\s+([0-9.]+) photons", out1 + out2)[-2:]
    assert len(nums) == 2 and nums[0] == nums[1], (nums, out2[-1500:])
    assert re.search(r"absorbed:.*3[0-8]\.[0-9]+%", out2), out2[-1500:]
    base = int(re.search(r"detected\s+([0-9]+) photons", out1).group(1))
    assert int(float(nums[0])) == base


def test_replay_jacobian_through_the_cli(tmp_path):
    """`-E seeds -O J`: the absorption Jacobian of the replayed photons, written by the reference's writer; its integral
    times the normaliser's inverse is sum_i w_i L_i, i.e. positive and finite, with one time gate"""
    rc, out1 = mcx(["--bench", "cube60", "-s", "rj", "-q", "1", "-S", "0", "-n", "1e5"], tmp_path)
    assert rc == 0
    rc, out2 = mcx(["--bench", "cube60", "-E", "rj_detp.jdat", "-O", "J", "-F", "mc2", "-s", "rjout", "-n", "1e5"], tmp_path)
    assert rc == 0, out2[-2000:]
    jac = np.fromfile(os.path.join(tmp_path, "rjout.mc2"), dtype=np.float32)
    assert jac.size == 216000 and np.isfinite(jac).all() and jac.min() >= 0 and jac.sum() > 0
    # normalised by unitinmm / sum(w): the integral is the weighted mean path length of the detected photons (mm)
    assert 20 < jac.astype(np.float64).sum() < 200


def test_detected_photon_flags_w(tmp_path):
    """test/testmcx.sh:134-136: `-w dspxvw -F jnii` writes six compressed blocks (detid, nscat, ppath, p, v, w0)"""
    rc, out = mcx(["--bench", "cube60", "-w", "dspxvw", "-F", "jnii", "-S", "0", "-n", "1e5"], tmp_path)
    assert rc == 0 and len(re.findall(r"compressing data \[zlib\]", out)) == 6, out[-1500:]


def test_progress_bar(tmp_path):
    """test/testmcx.sh:138-140: `-D P` ends with 'Progress: [...] 100%'"""
    rc, out = mcx(["--bench", "cube60", "-D", "P", "-n", "2e7", "-S", "0", "-d", "0"], tmp_path)
    assert rc == 0 and re.search(r"Progress: .* 100%", out), out[-800:]
    shown = sorted(set(int(x) for x in re.findall(r"\]\s+([0-9]+)%", out)))
    assert shown[-1] == 100 and len(shown) > 2           # intermediate values were drawn while the kernel ran


def test_unsupported_modes_fail_loudly(tmp_path):
    rc, out = mcx(["--bench", "cube60", "-n", "1e4", "-r", "-2", "-S", "0"], tmp_path)
    assert rc != 0 and "respin" in out


def test_repetitions_through_the_cli(tmp_path):
    """`-r 3`: the photon budget in three batches, each with the next slice of the seed stream, accumulated on the device
    and normalised once (mcxb200.h, mcxb_config.respin): same answer as one batch"""
    rc, out = mcx(["--bench", "cube60b", "-n", "300001", "-r", "3", "-F", "mc2", "-s", "rep", "-w", "DP"], tmp_path)
    assert rc == 0, out[-2000:]
    assert "repeat x3" in out and re.search(r"total simulated energy: 300001\.00\s+absorbed:\s*27\.[0-9]+%", out), out[-1500:]
    assert len(re.findall(r"simulation run#\s*[0-9]+", out)) == 3
    one = engine.run(benchmarks.get("cube60b", 300001))
    mc2 = np.fromfile(os.path.join(tmp_path, "rep.mc2"), dtype=np.float32).astype(np.float64)
    np.testing.assert_allclose(mc2.sum(), one["flux"].astype(np.float64).sum(), rtol=0.016)          # 3e5 packets each: sigma of the ratio 0.3 %
    raw = open(os.path.join(tmp_path, "rep.mch"), "rb").read()
    magic, version, maxmedia, detnum, colcount, totalphoton, detected, savedphoton = struct.unpack("<4s7I", raw[:32])
    respin = struct.unpack("<i", raw[44:48])[0]
    assert magic == b"MCXH" and totalphoton == 300001 and respin == 3
    assert abs(savedphoton - one["stat"]["detected"]) < 6 * np.sqrt(2.0 * savedphoton)


def test_trajectory_file_through_the_cli(tmp_path):
    """`-D M`: <session>.mct = 64-byte History header + {photon id, x, y, z, weight, source id} records
    (src/mcx_host.cpp:1666-1672, mcx_savedetphoton with his.detected == 0)"""
    rc, out = mcx(["--bench", "cube60", "-n", "2000", "-D", "M", "-F", "mc2", "-s", "tr", "-S", "0", "-d", "0"], tmp_path)
    assert rc == 0, out[-2000:]
    m = re.search(r"saved trajectory positions: ([0-9]+)", out)
    assert m, out[-1500:]
    raw = open(os.path.join(tmp_path, "tr.mct"), "rb").read()
    magic, version, maxmedia, detnum, colcount, totalphoton, detected, savedphoton = struct.unpack("<4s7I", raw[:32])
    assert magic == b"MCXH" and colcount == 6 and totalphoton == 2000 and detected == 0 and savedphoton == int(m.group(1))
    rec = np.frombuffer(raw[64:64 + 4 * 6 * savedphoton], dtype=np.float32).reshape(-1, 6)
    ids = rec[:, 0].copy().view(np.uint32)
    assert len(np.unique(ids)) == 2000 and ids.min() == 1 and ids.max() == 2000
    assert 60 < savedphoton / 2000.0 < 110                   # launch + ~80 scattering sites + end per packet


def test_mc2_and_mch_files_match_the_engine(tmp_path):
    """the reference's writers (mcx_savedata / mcx_savedetphoton, src/mcx_utils.c:886-998) fed by the CUDA engine:
    .mc2 = raw float32 fluence, .mch = 64-byte History header + detected-photon records"""
    n = 200000
    rc, out = mcx(["--bench", "cube60b", "-n", str(n), "-F", "mc2", "-s", "cli60b", "-w", "DP"], tmp_path)
    assert rc == 0, out[-2000:]
    mc2 = np.fromfile(os.path.join(tmp_path, "cli60b.mc2"), dtype=np.float32)
    assert mc2.size == 216000
    r = engine.run(benchmarks.get("cube60b", n))
    # same deck, same seed, same scheduler: the two runs are statistically equivalent (dynamic scheduling is not
    # stream-reproducible), compare integrals and the depth profile
    a, b = mc2.astype(np.float64).reshape(60, 60, 60), r["flux"][..., 0].astype(np.float64).transpose(2, 1, 0)
    np.testing.assert_allclose(a.sum(), b.sum(), rtol=0.01)
    np.testing.assert_allclose(a.sum(axis=(1, 2))[:30], b.sum(axis=(1, 2))[:30], rtol=0.05)
    raw = open(os.path.join(tmp_path, "cli60b.mch"), "rb").read()
    magic, version, maxmedia, detnum, colcount, totalphoton, detected, savedphoton = struct.unpack("<4s7I", raw[:32])
    unitinmm, seedbyte, normalizer = struct.unpack("<fIf", raw[32:44])
    assert magic == b"MCXH" and version == 1 and maxmedia == 2 and detnum == 4 and colcount == 3
    assert totalphoton == n and detected == savedphoton and unitinmm == 1.0 and seedbyte == 0
    rec = np.frombuffer(raw[64:64 + 4 * colcount * savedphoton], dtype=np.float32).reshape(-1, colcount)
    assert rec.shape[0] == savedphoton and set(np.unique(rec[:, 0]).astype(int)) == {1, 2, 3, 4}
    assert abs(savedphoton - r["stat"]["detected"]) < 6 * np.sqrt(2 * savedphoton)
    assert abs(rec[:, 1].mean() - r["detp"][1].mean()) < 0.05 * rec[:, 1].mean()
    assert normalizer == pytest.approx(r["stat"]["normalizer"], rel=1e-3)
    assert abs(absorbed(out) / 100 - r["stat"]["absorbed"]) < 0.005


def test_jnii_output_and_json_input_file(tmp_path):
    """a JSON deck read by the reference's parser (mcx_loadjson) and a .jnii volume written by mcx_savejnii"""
    import base64
    import zlib
    deck = {"Session": {"ID": "deck", "Photons": 50000, "RNGSeed": 1648335518, "DoMismatch": True},
            "Forward": {"T0": 0, "T1": 2e-9, "Dt": 1e-9},
            "Domain": {"OriginType": 1, "LengthUnit": 1, "Media": [{"mua": 0, "mus": 0, "g": 1, "n": 1}, {"mua": 0.005, "mus": 1, "g": 0.01, "n": 1.37}], "Dim": [40, 40, 30]},
            "Optode": {"Source": {"Type": "pencil", "Pos": [19, 19, 0], "Dir": [0, 0, 1]}},
            "Shapes": [{"Grid": {"Tag": 1, "Size": [40, 40, 30]}}]}
    with open(os.path.join(tmp_path, "deck.json"), "w") as f:
        json.dump(deck, f)
    rc, out = mcx(["-f", "deck.json", "-F", "jnii", "-d", "0"], tmp_path)
    assert rc == 0, out[-2000:]
    j = json.load(open(os.path.join(tmp_path, "deck.jnii")))
    nd = j["NIFTIData"]
    assert nd["_ArraySize_"][:4] == [40, 40, 30, 2]
    vol = np.frombuffer(zlib.decompress(base64.b64decode(nd["_ArrayZipData_"])), dtype=np.float32)
    assert vol.size == 40 * 40 * 30 * 2
    cfg = dict(nphoton=50000, vol=np.ones((40, 40, 30), np.uint8), prop=[[0, 0, 1, 1], [0.005, 1, 0.01, 1.37]], tstart=0, tend=2e-9, tstep=1e-9,
               srcpos=[19, 19, 0], srcdir=[0, 0, 1], issrcfrom0=1, isreflect=1, issavedet=0, seed=1648335518)
    r = engine.run(cfg)
    assert abs(absorbed(out) / 100 - r["stat"]["absorbed"]) < 0.01
    np.testing.assert_allclose(vol.astype(np.float64).sum(), r["flux"].astype(np.float64).sum(), rtol=0.04)        # 5e4 packets each: sigma of the ratio 0.8 %


# ------------------------------------------------------------------------------------------------ BASELINE config 1
QTEST_INP = """1000000              # total photon (not used)
29012392             # RNG seed, negative to generate
30.0 30.0 1.0        # source position (mm)
0 0 1                # initial directional vector
0.e+00 5.e-09 5.e-9  # time-gates(s): start, end, step
cubic60.json         # volume ('uchar' format)
1 60 10 50            # x: voxel size, dim, start/end indices
1 60 10 50            # y: voxel size, dim, start/end indices
1 60 1  20            # z: voxel size, dim, start/end indices
1                    # num of media
1 0.01 0.005 1.0  # scat(1/mm), g, mua (1/mm), n
4\t1            # detector number and radius (mm)
30.0\t20.0\t1.0  # detector 1 position (mm)
30.0\t40.0\t1.0  # ...
20.0\t30.0\t1.0
40.0\t30.0\t1.0
"""
QTEST_SHAPES = {"Shapes": [{"Name": "cube60"}, {"Origin": [0, 0, 0]}, {"Grid": {"Tag": 1, "Size": [60, 60, 60]}}]}


def test_quicktest_inp_deck_through_the_cli(tmp_path):
    """BASELINE config 1: the legacy .inp deck of example/quicktest (qtest.inp:1-16 + cubic60.json, command of
    run_qtest.sh:3 at 1e6 photons) read by the reference's own mcx_loadconfig (src/mcx_utils.c:2087-2398) and run on
    the CUDA engine; compared with the engine driven through the Python mirror of the same deck (benchmarks.qtest)
    and with the committed reference series of cube60 (same geometry; n = 1 instead of 1.37 changes nothing with
    matched boundaries: the medium-1 row is the only difference and there is no index mismatch in either)."""
    with open(os.path.join(tmp_path, "qtest.inp"), "w") as f:
        f.write(QTEST_INP)
    with open(os.path.join(tmp_path, "cubic60.json"), "w") as f:
        json.dump(QTEST_SHAPES, f)
    n = 1000000
    rc, out = mcx(["-A", "-n", str(n), "-f", "qtest.inp", "-F", "mc2", "-s", "qt", "-w", "DP"], tmp_path)
    assert rc == 0, out[-2000:]
    r = engine.run(benchmarks.get("qtest", n))
    a = absorbed(out) / 100
    assert abs(a - r["stat"]["absorbed"]) < 0.003, (a, r["stat"]["absorbed"])
    m = re.search(r"detected\s+([0-9]+) photons", out)
    assert m and abs(int(m.group(1)) - r["stat"]["detected"]) < 6 * np.sqrt(2.0 * r["stat"]["detected"])
    mc2 = np.fromfile(os.path.join(tmp_path, "qt.mc2"), dtype=np.float32)
    assert mc2.size == 216000
    x, y = mc2.astype(np.float64).reshape(60, 60, 60), r["flux"][..., 0].astype(np.float64).transpose(2, 1, 0)
    np.testing.assert_allclose(x.sum(), y.sum(), rtol=0.01)
    np.testing.assert_allclose(x.sum(axis=(1, 2))[:30], y.sum(axis=(1, 2))[:30], rtol=0.03)
    # the source voxel (29,29,0) after the 1-based -> 0-based shift of the .inp reader holds the peak
    assert np.unravel_index(np.argmax(x), x.shape) == (0, 29, 29)
    raw = open(os.path.join(tmp_path, "qt.mch"), "rb").read()
    magic, version, maxmedia, detnum, colcount, totalphoton, detected, savedphoton = struct.unpack("<4s7I", raw[:32])
    assert magic == b"MCXH" and maxmedia == 1 and detnum == 4 and colcount == 2 and totalphoton == n
    rec = np.frombuffer(raw[64:64 + 4 * colcount * savedphoton], dtype=np.float32).reshape(-1, colcount)
    assert set(np.unique(rec[:, 0]).astype(int)) == {1, 2, 3, 4}


# ------------------------------------------------------------------------------------------------ output formats read back
def _jdata_array(node):
    """text-JData annotated array (zlib + base64) -> numpy, in the annotated shape"""
    import base64
    import zlib
    dt = {"single": "<f4", "uint32": "<u4", "int32": "<i4", "uint8": "u1", "double": "<f8"}[node["_ArrayType_"]]
    return np.frombuffer(zlib.decompress(base64.b64decode(node["_ArrayZipData_"])), dtype=dt).reshape(node["_ArraySize_"])


def test_bnii_tx3_and_nii_volumes_read_back(tmp_path):
    """the remaining volume formats of mcx_savedata (src/mcx_utils.c:886-951) fed by the CUDA engine:
      .bnii  binary JData written by mcx_savebnii (:598-735): NIFTIHeader.Dim = [Nx,Ny,Nz,Ngate,Nsrc], float32 payload
      .tx3   GL_RGBA32F tag + 3 ints + raw floats (:944-950)
      .nii   348-byte NIfTI-1 header + 4-byte extender + float32 data (:487-585).  HAZARD kept from the reference: for
             float data mcx_savenii writes its freshly malloc'ed `logval` buffer, never `dat` (:507-510) -- the header is
             right, the voxel data are whatever malloc returned; only the header is asserted here."""
    import bjdata
    n = 200000
    ref = engine.run(dict(benchmarks.get("cube60b", n), tend=2e-9, tstep=1e-9))       # two gates
    want = ref["flux"].astype(np.float64)                                             # (60,60,60,2)
    common = ["--bench", "cube60b", "-n", str(n), "-d", "0", "--json", '{"Forward":{"T0":0,"T1":2e-9,"Dt":1e-9}}']

    rc, out = mcx(common + ["-F", "bnii", "-s", "vb"], tmp_path)
    assert rc == 0, out[-2000:]
    d = bjdata.load(os.path.join(tmp_path, "vb.bnii"))
    hdr = d["NIFTIHeader"]
    assert list(hdr["Dim"]) == [60, 60, 60, 2, 1] and hdr["DataType"] == "single" and hdr["NIIHeaderSize"] == 348
    assert hdr["VoxelSize"][0] == 1.0 and hdr["VoxelSize"][3] == pytest.approx(1e-9) and hdr["Name"] == "vb"
    assert "Fluence rate" in hdr["Description"]
    vol = bjdata.decode_array(d["NIFTIData"]).astype(np.float64)
    assert vol.size == 60 * 60 * 60 * 2
    gates = vol.reshape(2, -1).sum(1)                                                 # x fastest ... gate slowest in memory
    np.testing.assert_allclose(gates, want.reshape(-1, 2, order="F").sum(0), rtol=0.02)
    np.testing.assert_allclose(vol.reshape(2, 60, 60, 60)[0].sum(axis=(1, 2))[:25], want[..., 0].sum(axis=(0, 1))[:25], rtol=0.06)

    rc, out = mcx(common + ["-F", "tx3", "-s", "vt"], tmp_path)
    assert rc == 0, out[-2000:]
    raw = open(os.path.join(tmp_path, "vt.tx3"), "rb").read()
    glformat, nx, ny, nz = struct.unpack("<I3i", raw[:16])
    assert glformat == 0x8814 and (nx, ny, nz) == (60, 60, 60)                        # GL_RGBA32F
    tx = np.frombuffer(raw[16:], dtype=np.float32).astype(np.float64)
    assert tx.size == 2 * 216000
    np.testing.assert_allclose(tx.reshape(2, -1).sum(1), gates, rtol=0.02)

    rc, out = mcx(common + ["-F", "nii", "-s", "vn"], tmp_path)
    assert rc == 0, out[-2000:]
    raw = open(os.path.join(tmp_path, "vn.nii"), "rb").read()
    assert len(raw) == 352 + 4 * 2 * 216000
    sizeof_hdr = struct.unpack("<i", raw[:4])[0]
    dim = struct.unpack("<8h", raw[40:56])
    datatype, bitpix = struct.unpack("<hh", raw[70:74])
    pixdim = struct.unpack("<8f", raw[76:108])
    vox_offset = struct.unpack("<f", raw[108:112])[0]
    assert sizeof_hdr == 348 and dim[:5] == (4, 60, 60, 60, 2) and datatype == 16 and bitpix == 32      # NIFTI_TYPE_FLOAT32
    assert pixdim[1:4] == (1.0, 1.0, 1.0) and pixdim[4] == pytest.approx(1e-9 * 1e6) and vox_offset == 352.0
    assert raw[344:348] == b"n+1\x00"


def test_detected_photon_jdat_fields_read_back(tmp_path):
    """`-w dspxvw` + the default jnii family: <session>_detp.jdat (mcx_savejdet, src/mcx_utils.c:1010-1182) holds one
    annotated array per record field; read back with the column semantics of utils/loadmch.m:64-75 (detid 1, nscat M,
    ppath M, mom M, p 3, v 3, w0 1) and compared with the same run made through the Python mirror."""
    n = 300000
    rc, out = mcx(["--bench", "cube60b", "-n", str(n), "-w", "dspxvw", "-F", "jnii", "-s", "jd", "-S", "0"], tmp_path)
    assert rc == 0, out[-2000:]
    j = json.load(open(os.path.join(tmp_path, "jd_detp.jdat")))
    info, pd = j["MCXData"]["Info"], j["MCXData"]["PhotonData"]
    assert info["DetNum"] == 4 and info["MediaNum"] == 2 and info["TotalPhoton"] == n and info["LengthUnit"] == 1
    ndet = info["DetectedPhoton"]
    assert info["SavedPhoton"] == ndet and set(pd) == {"detid", "nscat", "ppath", "p", "v", "w0"}
    detid, nscat, ppath, pos, vel, w0 = (_jdata_array(pd[k]) for k in ("detid", "nscat", "ppath", "p", "v", "w0"))
    assert detid.shape == (ndet, 1) and nscat.shape == (ndet, 2) and ppath.shape == (ndet, 2)
    assert pos.shape == (ndet, 3) and vel.shape == (ndet, 3) and w0.shape == (ndet, 1)
    assert set(np.unique(detid)) == {1, 2, 3, 4}
    # nscat: the kernel keeps the counts as uint32 BIT patterns in the float record (src/mcx_core.cl:2515, here too) and
    # mcx_savejdet converts the float VALUE to uint (src/mcx_utils.c:1090) -- denormals, i.e. zeros, in the reference's own
    # files as well; pmcxcl / .mch users reinterpret the bits instead (checked below on the Python side)
    assert (nscat == 0).all() and (ppath[:, 0] > 0).all() and (ppath[:, 1] == 0).all()                # medium 2 is never entered
    np.testing.assert_allclose(np.linalg.norm(vel, axis=1), 1.0, atol=1e-4)
    assert (vel[:, 2] < 0).all() and (pos[:, 2] <= 1e-3).all() and (w0 == 1).all()                    # leaving through z = 0
    r = engine.run(dict(benchmarks.get("cube60b", n), savedetflag="dspxvw"))
    d = r["detp"]                                                       # (reclen, ndet): detid, nscat x2, ppath x2, p, v, w0
    assert d.shape[0] == 12 and abs(d.shape[1] - ndet) < 6 * np.sqrt(2.0 * ndet)
    assert abs(ppath[:, 0].mean() - d[3].mean()) < 0.05 * d[3].mean()
    counts = np.ascontiguousarray(d[1]).view(np.uint32).astype(np.float64)
    assert counts.min() >= 1 and abs(counts.mean() / ppath[:, 0].mean() - 1.0) < 0.1                 # mus = 1 per voxel: one event per unit path


# ---- the extended modes through the unchanged CLI (JSON keys of src/mcx_utils.c:2510-2563, 2944-2949, 3049-3073) ----------------
POLARIZED_DECK = """{
 "Session": {"ID": "onelayer", "DoMismatch": 0, "DoAutoThread": 1, "MaxDetPhoton": 1000000, "BCFlags": "______001000",
             "SaveDataMask": "IXVSPW", "Photons": 200000, "MinEnergy": 0.01},
 "Forward": {"T0": 0, "T1": 5e-09, "Dt": 5e-09},
 "Optode": {"Source": {"Pos": [10, 10, 0], "Dir": [0, 0, 1], "IQUV": [1, 1, 0, 0], "WaveLength": 632.8}},
 "Domain": {"OriginType": 1, "LengthUnit": 1, "Media": [{"mua": 0, "mus": 0, "g": 1, "n": 1}, {"mua": 0, "mus": 0, "g": 1, "n": 1}],
            "MieScatter": [{"mua": 0.0, "radius": 1.015, "rho": 0.0001152, "nsph": 1.59, "nmed": 1.33}],
            "MediaFormat": "byte", "Dim": [20, 20, 10]},
 "Shapes": [{"Grid": {"Tag": 1, "Size": [20, 20, 10]}}]
}"""


def test_polarised_deck_of_the_reference_examples(tmp_path):
    """example/polarized/onelayer.json (the volume written as a Grid shape instead of the zipped array): the CLI's own Mie
    code fills the media table and the Mueller matrices (mcx_prep_polarized), the engine tracks Stokes vectors, the detected
    records carry them (`I` of SaveDataMask) and reach the unchanged .jdat writer"""
    (tmp_path / "onelayer.json").write_text(POLARIZED_DECK)
    rc, out = mcx(["-f", "onelayer.json", "-F", "jnii", "-S", "0"], tmp_path)
    assert rc == 0, out[-2000:]
    m = re.search(r"detected (\d+) photons", out)
    assert m and int(m.group(1)) > 20000, out[-1500:]                   # the +z face is a detector (BCFlags); mua = 0: most packets leave there
    assert absorbed(out) < 0.01
    jd = json.loads((tmp_path / "onelayer_detp.jdat").read_text())
    info = jd["MCXData"]["Info"]
    assert info["DetectedPhoton"] == int(m.group(1)) and info["SavedPhoton"] == int(m.group(1))
    # the reference's .jdat writer knows seven record fields and no Stokes columns (src/mcx_utils.c:1074): what it can name is there
    assert {"nscat", "ppath", "p", "v", "w0"} <= set(jd["MCXData"]["PhotonData"])


def test_rf_frequency_through_the_cli(tmp_path):
    """Optode.Source.Frequency (src/mcx_utils.c:2944-2945) turns the run into an RF forward run: same packets, same absorbed
    fraction.  (The adjoint types cannot be reached from the CLI: its JSON reader takes a detector's position from the
    Detector ARRAY whenever the detector object has a third member such as "Dir" (src/mcx_utils.c:3047-3059), which
    dereferences a NULL `next` for one detector -- `mcxcl --dumpjson` with such a deck dies before any device is touched.
    The adjoint path is tested through pmcxcl, tests/test_gpu_pmcxcl.py.)"""
    rc, out = mcx(["--bench", "cube60b", "-n", "2e5", "-S", "0", "--json", '{"Optode":{"Source":{"Frequency":1e8}}}'], tmp_path)
    assert rc == 0 and re.search(r"absorbed:.*27\.[0-9]+%", out), out[-1500:]
