"""Host-side logic of the multi-GPU path on CPU: photon split, seed slices, gather plan, and the collective step
itself (mcxcl_b200.multigpu.combine_tensors) run by two gloo processes over 127.0.0.1 on CPU tensors."""
import os
import socket

import numpy as np
import pytest

from mcxcl_b200 import multigpu


def test_split_photons_equal_and_weighted():
    assert multigpu.split_photons(10, [1, 1, 1]) == [4, 3, 3]
    assert multigpu.split_photons(1000000000, [1] * 8) == [125000000] * 8
    s = multigpu.split_photons(1000, [3, 1])
    assert s == [750, 250] and sum(s) == 1000
    s = multigpu.split_photons(7, [0, 1, 1])
    assert s[0] == 0 and sum(s) == 7                     # a zero weight never receives the remainder
    assert multigpu.split_photons(0, [1, 1]) == [0, 0]
    for bad in ([], [0, 0], [-1, 2]):
        with pytest.raises(ValueError):
            multigpu.split_photons(10, bad)


def test_gather_plan_clips_at_the_buffer_size():
    assert multigpu.gather_plan([3, 0, 5], 100) == ([0, 3, 3], [3, 0, 5], 8)
    assert multigpu.gather_plan([60, 60, 60], 100) == ([0, 60, 100], [60, 40, 0], 100)
    assert multigpu.gather_plan([], 10) == ([], [], 0)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, case, outdir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(100 + rank)
        fieldlen, reclen, maxdet = 5000, 3, case["maxdet"]
        field = torch.from_numpy(rs.uniform(0, 1, fieldlen).astype(np.float32))
        energy = torch.tensor([10.0 + rank, 100.0 * (rank + 1)], dtype=torch.float64)
        count = case["counts"][rank]
        rec = np.zeros((maxdet, reclen), np.float32)
        stored = min(count, maxdet)
        rec[:stored, 0] = rank + 1
        rec[:stored, 1] = np.arange(stored)
        seeds = np.zeros((maxdet, 2), np.int64)
        seeds[:stored, 0] = rank
        seeds[:stored, 1] = np.arange(stored)
        records = torch.from_numpy(rec.ravel()) if case["savedet"] else None
        np.save(os.path.join(outdir, "field%d.npy" % rank), field.numpy().copy())
        # seed slices: exclusive prefix sum of unequal per-rank thread counts
        offs = multigpu.seed_offsets(dist, world, 1000 * (rank + 1))
        assert offs == [0, 1000], offs
        out, oseeds, counts = multigpu.combine_tensors(dist, rank, world, field, energy, count, records, reclen, maxdet,
                                                       torch.from_numpy(seeds.ravel()) if case["seeds"] else None)
        if rank == 0:
            np.save(os.path.join(outdir, "sum.npy"), field.numpy())
            np.save(os.path.join(outdir, "energy.npy"), energy.numpy())
            np.save(os.path.join(outdir, "counts.npy"), np.array(counts))
            if out is not None:
                np.save(os.path.join(outdir, "rec.npy"), out.numpy().reshape(-1, reclen))
            if oseeds is not None:
                np.save(os.path.join(outdir, "seeds.npy"), oseeds.numpy().reshape(-1, 2))
        else:
            assert out is None and oseeds is None
    finally:
        dist.destroy_process_group()


CASES = {
    "ragged": dict(counts=[7, 3], maxdet=50, savedet=True, seeds=True),
    "empty_rank": dict(counts=[0, 9], maxdet=50, savedet=True, seeds=False),
    "overflow": dict(counts=[40, 80], maxdet=50, savedet=True, seeds=True),
    "no_detectors": dict(counts=[0, 0], maxdet=0, savedet=False, seeds=False),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_combine_tensors_world_size_2_gloo(tmp_path, name):
    import torch.multiprocessing as mp
    case = CASES[name]
    mp.spawn(_worker, args=(2, _free_port(), case, str(tmp_path)), nprocs=2, join=True)
    want = np.load(tmp_path / "field0.npy") + np.load(tmp_path / "field1.npy")
    np.testing.assert_array_equal(np.load(tmp_path / "sum.npy"), want)
    np.testing.assert_array_equal(np.load(tmp_path / "energy.npy"), [21.0, 300.0])
    np.testing.assert_array_equal(np.load(tmp_path / "counts.npy"), case["counts"])
    if not case["savedet"]:
        assert not (tmp_path / "rec.npy").exists()
        return
    rec = np.load(tmp_path / "rec.npy")
    n0 = min(case["counts"][0], case["maxdet"])
    n1 = min(case["counts"][1], case["maxdet"] - n0)
    assert rec.shape == (n0 + n1, 3)
    assert (rec[:n0, 0] == 1).all() and (rec[n0:, 0] == 2).all()
    np.testing.assert_array_equal(rec[:n0, 1], np.arange(n0))
    np.testing.assert_array_equal(rec[n0:, 1], np.arange(n1))
    if case["seeds"]:
        sd = np.load(tmp_path / "seeds.npy")
        assert sd.shape == (n0 + n1, 2) and (sd[:n0, 0] == 0).all() and (sd[n0:, 0] == 1).all()
        np.testing.assert_array_equal(sd[n0:, 1], np.arange(n1))


def test_rank_seed_slices_are_disjoint_parts_of_one_stream(lib):
    """rank r seeds its nthread streams from records [r*nthread, (r+1)*nthread) of ONE rand() stream
    (src/mcx_host.cpp:759-768): no two ranks share a stream"""
    nthread, world = 1024, 4
    seen = set()
    for r in range(world):
        part = np.zeros(4 * nthread, dtype=np.uint32)
        lib.mcxb_fill_seeds(1648335518, r * nthread, nthread, part.ctypes.data)
        rows = {tuple(x) for x in part.reshape(-1, 4).tolist()}
        assert len(rows) == nthread and not (rows & seen)
        seen |= rows
