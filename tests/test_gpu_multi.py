"""Multi-GPU paths on a box with at least two B200s (skipped on a single-GPU box):
  * one process per GPU over NCCL (mcxcl_b200.multigpu.run_distributed, launched with torch.distributed.run);
  * several GPUs from ONE process behind the C ABI (mcxb_run_simulation_multi: NCCL reduce + record gather inside the
    library), called directly and through the reference boundary (`mcxcl -G 11`, integration/mcx_cuda_host.cpp)."""
import json
import os
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

from mcxcl_b200 import benchmarks, engine, hostcfg

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import json, os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from mcxcl_b200 import benchmarks, multigpu
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = benchmarks.get("cube60b", 400001)
cfg["issaveseed"] = 1
res = multigpu.run_distributed(cfg, workload=[3.0, 1.0])
if dist.get_rank() == 0:
    raw = res["field"].astype(np.float64) / res["normalizer"]
    ids = res["detp"][:, 0].astype(int)
    print("RESULT " + json.dumps(dict(energytot=res["energytot"], absorbed=res["absorbed"], detected=res["detected"], saved=int(res["saved"]),
          shares=res["shares"], rawsum=float(raw.sum()), energyabs=res["energyabs"], ndet=[int((ids == k).sum()) for k in (1, 2, 3, 4)],
          nseeds=int(res["seeds"].shape[0]), uniqseeds=int(len({tuple(x) for x in res["seeds"].tolist()})))), flush=True)
dist.barrier()
dist.destroy_process_group()
'''


def need_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")


def test_nccl_photon_shards_combine_on_rank0(tmp_path):
    need_two_gpus()
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", str(script)], capture_output=True, text=True, timeout=600)
    m = re.search(r"RESULT (\{.*\})", r.stdout)
    assert m, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.loads(m.group(1))
    assert res["shares"] == [300001, 100000]                        # workload 3:1, remainder to the first rank
    assert res["energytot"] == 400001                               # every packet launched exactly once over the two ranks
    one = engine.run(benchmarks.get("cube60b", 400001))["stat"]
    assert abs(res["absorbed"] - one["absorbed"]) < 0.008        # two runs of 4e5 packets: sigma of the difference 1.5e-3
    assert abs(res["detected"] - one["detected"]) < 6 * np.sqrt(2 * one["detected"])
    assert res["saved"] == res["detected"] == sum(res["ndet"]) == res["nseeds"] == res["uniqseeds"]
    assert res["rawsum"] * 0.005 == pytest.approx(res["energyabs"], rel=2e-3)      # sum(field)*mua == absorbed energy


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_multi_gpu_call_behind_the_c_abi(exchange, monkeypatch):
    """mcxb_run_simulation_multi on two devices: workload split, disjoint seed slices, then the exchange -- by default ONE
    kernel on device 0 that sums the peers' volumes over NVLink-mapped memory plus peer-to-peer copies of the
    variable-length records and their RNG states; with MCXB_MULTI_EXCHANGE=nccl an NCCL reduce / all-gather / send-recv"""
    need_two_gpus()
    if exchange == "nccl":
        monkeypatch.setenv("MCXB_MULTI_EXCHANGE", "nccl")
    else:
        monkeypatch.delenv("MCXB_MULTI_EXCHANGE", raising=False)
    cfg = benchmarks.get("cube60b", 400001)
    cfg["issaveseed"] = 1
    p = hostcfg.prepare(cfg)
    r = engine.run_prepared_multi(p, [0, 1], workload=[3.0, 1.0])
    m = r["multi"]
    assert m["ndev"] == 2 and (m["nccl_version"] >= 20000 if exchange == "nccl" else m["nccl_version"] == 0)
    assert m["share"] == [300001, 100000]
    assert r["energytot"] == 400001                                 # every packet launched exactly once over the two devices
    assert r["detected"] == sum(m["detected"]) == r["saved"] == r["detp"].shape[0] == r["seeds"].shape[0]
    assert min(m["detected"]) > 0 and abs(m["detected"][0] / m["detected"][1] - 3.0) < 0.6
    assert len({tuple(x) for x in r["seeds"].tolist()}) == r["saved"]        # no RNG state twice: the seed slices are disjoint
    assert set(np.unique(r["detp"][:, 0]).astype(int)) == {1, 2, 3, 4}
    raw = r["field"].astype(np.float64) / r["normalizer"]
    assert raw.sum() * 0.005 == pytest.approx(r["energyabs"], rel=2e-3)      # sum(field) * mua == absorbed energy
    one = engine.run_prepared(p)
    assert abs(r["absorbed"] - one["absorbed"]) < 0.008
    assert abs(r["detected"] - one["detected"]) < 6 * np.sqrt(2 * one["detected"])
    assert r["normalizer"] == pytest.approx(one["normalizer"], rel=1e-6)
    # equal shares by default, records clipped at maxdetphoton like the reference (src/mcx_host.cpp:1207-1216)
    p2 = hostcfg.prepare(dict(benchmarks.get("cube60b", 400000), maxdetphoton=1000))
    r2 = engine.run_prepared_multi(p2, [0, 1])
    assert r2["multi"]["share"] == [200000, 200000] and r2["detected"] > 1000 and r2["saved"] == 1000 == r2["detp"].shape[0]
    # a replay is single-device only (src/mcx_host.cpp:723)
    with pytest.raises(RuntimeError, match="single device"):
        rp = hostcfg.prepare(dict(benchmarks.get("cube60b", r["saved"]), seed=np.ascontiguousarray(r["seeds"]).view(np.uint8).reshape(-1, 16).T.copy(),
                                  detphotons=np.ascontiguousarray(r["detp"].T), outputtype="jacobian"))
        engine.run_prepared_multi(rp, [0, 1])


def test_two_devices_from_one_process_through_the_reference_boundary(tmp_path):
    need_two_gpus()
    exe = os.path.join(ROOT, "integration", "_build", "mcxcl")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/mcxcl not built")
    r = subprocess.run([exe, "--bench", "cube60b", "-n", "400001", "-G", "11", "-W", "3,1", "-w", "DP", "-F", "mc2", "-s", "two"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout + r.stderr)
    assert r.returncode == 0, out[-2000:]
    assert "with 2 devices" in out and re.search(r"total simulated energy: 400001\.00\s+absorbed:\s*27\.[0-9]+%", out), out[-1500:]
    assert re.search(r"np=300001\.0", out) and re.search(r"np=100000\.0", out)
    assert "devices combined over" in out
    # the records of BOTH devices reach the .mch file the reference's writer produces
    m = re.search(r"detected\s+([0-9]+) photons", out)
    raw = open(os.path.join(tmp_path, "two.mch"), "rb").read()
    magic, version, maxmedia, detnum, colcount, totalphoton, detected, savedphoton = struct.unpack("<4s7I", raw[:32])
    assert magic == b"MCXH" and colcount == 3 and totalphoton == 400001 and savedphoton == detected == int(m.group(1))
    rec = np.frombuffer(raw[64:64 + 4 * colcount * savedphoton], dtype=np.float32).reshape(-1, colcount)
    assert set(np.unique(rec[:, 0]).astype(int)) == {1, 2, 3, 4} and 1500 < savedphoton < 2200
    mc2 = np.fromfile(os.path.join(tmp_path, "two.mc2"), dtype=np.float32)
    one = engine.run(benchmarks.get("cube60b", 400001))
    np.testing.assert_allclose(mc2.astype(np.float64).sum(), one["flux"].astype(np.float64).sum(), rtol=0.015)
