"""Multi-GPU paths on a box with at least two B200s (skipped on a single-GPU box):
  * one process per GPU over NCCL (mcxcl_b200.multigpu.run_distributed, launched with torch.distributed.run);
  * several GPUs from one process through the reference boundary (`mcxcl -G 11`, integration/mcx_cuda_host.cpp)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from mcxcl_b200 import benchmarks, engine

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
    assert abs(res["absorbed"] - one["absorbed"]) < 0.004
    assert abs(res["detected"] - one["detected"]) < 6 * np.sqrt(2 * one["detected"])
    assert res["saved"] == res["detected"] == sum(res["ndet"]) == res["nseeds"] == res["uniqseeds"]
    assert res["rawsum"] * 0.005 == pytest.approx(res["energyabs"], rel=2e-3)      # sum(field)*mua == absorbed energy


def test_two_devices_from_one_process_through_the_reference_boundary(tmp_path):
    need_two_gpus()
    exe = os.path.join(ROOT, "integration", "_build", "mcxcl")
    if not os.path.exists(exe):
        pytest.skip("integration/_build/mcxcl not built")
    r = subprocess.run([exe, "--bench", "cube60b", "-n", "400001", "-G", "11", "-W", "3,1", "-S", "0"], cwd=tmp_path, capture_output=True, text=True, timeout=600)
    out = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout + r.stderr)
    assert r.returncode == 0, out[-2000:]
    assert "with 2 devices" in out and re.search(r"total simulated energy: 400001\.00\s+absorbed:\s*27\.[0-9]+%", out), out[-1500:]
    assert re.search(r"np=300001\.0", out) and re.search(r"np=100000\.0", out)
