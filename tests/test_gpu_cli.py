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
    rc, out = mcx(["--bench", "cube60", "-n", "1e4", "-r", "2", "-S", "0"], tmp_path)
    assert rc != 0 and "respin" in out


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
    np.testing.assert_allclose(vol.astype(np.float64).sum(), r["flux"].astype(np.float64).sum(), rtol=0.02)


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
