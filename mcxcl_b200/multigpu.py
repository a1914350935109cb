"""Photon sharding over the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

The reference splits the photon budget over OpenCL devices by workload weight (``-G`` mask, ``-W`` weights;
src/mcx_host.cpp:650-662, 1011-1012, 1098-1104), gives device *i* the next slice of ONE rand() stream
(:759-768), lets every device fill a private fluence volume and sums volumes, energies and detected-photon
lists on the host (:1218-1232, 1292-1306).  Here every rank runs the persistent kernel on its share with its
seed slice, then ONE collective step combines the results on the device before normalisation:

    reduce(SUM, float32 volume) -> rank 0      (NCCL over NVLink / NVSwitch)
    reduce(SUM, float64 {escaped, launched})   -> rank 0
    all_gather(detected counts) + variable-length send/recv of the detected-photon records -> rank 0

`split_photons` / `gather_plan` are pure host logic (tested with gloo on CPU); `run_distributed` needs GPUs.
"""
import numpy as np

from . import engine, hostcfg


def split_photons(nphoton, workload):
    """Per-rank photon counts: nphoton*w_i/sum(w) rounded down, the remainder going to the first ranks
    (the reference computes threadphoton/oddphoton per device from the same ratio, src/mcx_host.cpp:1011-1012)."""
    w = np.asarray(workload, dtype=np.float64)
    if w.ndim != 1 or w.size == 0 or (w < 0).any() or w.sum() <= 0:
        raise ValueError("workload must be a non-empty list of non-negative weights with a positive sum")
    share = np.floor(int(nphoton) * (w / w.sum())).astype(np.int64)
    rem = int(nphoton) - int(share.sum())
    order = [i for i in range(w.size) if w[i] > 0]
    for k in range(rem):
        share[order[k % len(order)]] += 1
    return [int(x) for x in share]


def gather_plan(counts, maxdetphoton):
    """offsets/lengths of each rank's detected-photon records in rank 0's buffer (exclusive scan, clipped to
    the buffer size the way the reference clips at maxdetphoton, src/mcx_host.cpp:1207-1216)."""
    offs, lens, at = [], [], 0
    for c in counts:
        n = max(0, min(int(c), int(maxdetphoton) - at))
        offs.append(at)
        lens.append(n)
        at += n
    return offs, lens, at


def seed_offsets(dist, world, nthread, device=None):
    """exclusive prefix sum over ranks of the number of RNG streams (= threads) each rank runs"""
    import torch
    mine = torch.tensor([int(nthread)], dtype=torch.int64, device=device)
    every = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, mine)
    counts = [int(x) for x in every.tolist()]
    return [sum(counts[:r]) for r in range(world)]


class _DevArray:
    """exposes a raw device pointer to torch through __cuda_array_interface__ (zero copy)"""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=2)


def _as_tensor(ptr, shape, typestr, device):
    import torch
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


def combine_tensors(dist, rank, world, field, energy, detcount, records, reclen, maxdetphoton, seeds=None):
    """The collective step on plain torch tensors (CUDA tensors over NCCL in production, CPU tensors over gloo
    in tests/test_multigpu_host.py):

        field    float32[fieldlen]   raw deposits of this rank          -> summed in place on rank 0
        energy   float64[2]          {escaped, launched}                -> summed in place on rank 0
        detcount int                 photons this rank detected (may exceed maxdetphoton)
        records  float32[>= stored*reclen]  this rank's detected-photon records
        seeds    int64[>= stored*2] or None  RNG states of the detected photons (issaveseed)

    Returns (records on rank 0 or None, seeds on rank 0 or None, per-rank detected counts)."""
    import torch
    dist.reduce(field, dst=0, op=dist.ReduceOp.SUM)
    dist.reduce(energy, dst=0, op=dist.ReduceOp.SUM)
    if records is None:
        return None, None, [0] * world
    mine = detcount.to(torch.int64).reshape(1) if torch.is_tensor(detcount) else torch.tensor([int(detcount)], dtype=torch.int64, device=field.device)
    gathered = torch.zeros(world, dtype=torch.int64, device=field.device)
    dist.all_gather_into_tensor(gathered, mine)
    counts = [int(x) for x in gathered.tolist()]          # ONE device->host sync for all ranks' counts
    reclen = max(1, int(reclen))
    stored = [min(x, int(maxdetphoton)) for x in counts]
    offs, lens, total = gather_plan(stored, maxdetphoton)
    ops, out, out_seeds = [], None, None
    if rank == 0:
        out = torch.empty(total * reclen, dtype=torch.float32, device=field.device)
        out[:lens[0] * reclen] = records[:lens[0] * reclen]
        if seeds is not None:
            out_seeds = torch.empty(total * 2, dtype=torch.int64, device=field.device)
            out_seeds[:lens[0] * 2] = seeds[:lens[0] * 2]
        for r in range(1, world):
            if lens[r]:
                ops.append(dist.P2POp(dist.irecv, out[offs[r] * reclen:(offs[r] + lens[r]) * reclen], r))
                if seeds is not None:
                    ops.append(dist.P2POp(dist.irecv, out_seeds[offs[r] * 2:(offs[r] + lens[r]) * 2], r))
    elif lens[rank]:
        ops.append(dist.P2POp(dist.isend, records[:lens[rank] * reclen].contiguous(), 0))
        if seeds is not None:
            ops.append(dist.P2POp(dist.isend, seeds[:lens[rank] * 2].contiguous(), 0))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return out, out_seeds, counts


def combine(sim, dist, rank, world, device):
    """finalize this rank's accumulators and run the collective step on the engine's device buffers (zero copy);
    returns (gathered detected records on rank 0 or None, gathered seeds or None, per-rank counts)"""
    sim.finalize()
    ptr = sim.devptrs()
    field = _as_tensor(ptr["field"], (sim.fieldlen,), "<f4", device)
    energy = _as_tensor(ptr["energy"], (2,), "<f8", device)
    c = sim.p.c
    if not (c.issavedet and ptr["detphoton"]):
        return combine_tensors(dist, rank, world, field, energy, 0, None, 0, 0)
    reclen = max(1, sim.reclen)
    mine = _as_tensor(ptr["detcount"], (1,), "<i4", device)      # stays on the device: gathered with everyone else's
    local = _as_tensor(ptr["detphoton"], (c.maxdetphoton * reclen,), "<f4", device)
    seeds = _as_tensor(ptr["seeddata"], (c.maxdetphoton * 2,), "<i8", device) if (c.issaveseed and ptr["seeddata"]) else None
    return combine_tensors(dist, rank, world, field, energy, mine, local, reclen, c.maxdetphoton, seeds)


def _to_host(t):
    """device tensor -> numpy through PINNED host memory: the gathered records of 8 GPUs are ~100 MB per call, which a
    pageable `.cpu()` moves at a few GB/s; torch's caching host allocator hands the same pinned block out again on the next
    call, and the returned array keeps it alive for as long as the caller holds it"""
    import torch
    if t.device.type != "cuda":
        return t.numpy()
    buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    buf.copy_(t)
    return buf.numpy()


def run_distributed(cfg, workload=None):
    """`pmcxcl.run`-style call executed by every rank of an initialised torch.distributed NCCL group; rank 0
    returns the combined result dictionary, the other ranks return None."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    device = torch.device("cuda", torch.cuda.current_device())
    if isinstance(cfg, hostcfg.Prepared):
        # already validated / pre-processed input (what a front-end holds after mcx_preprocess): only the budget changes
        total = int(cfg.c.nphoton)
        shares = split_photons(total, workload or [1.0] * world)
        p = cfg.clone_for(nphoton=shares[rank])
    else:
        total = int(cfg["nphoton"])
        shares = split_photons(total, workload or [1.0] * world)
        p = hostcfg.prepare(dict(cfg, nphoton=shares[rank]))
    with engine.Simulation(p, device.index) as sim:
        # rank r takes the slice of the ONE rand() stream that follows the slices of ranks 0..r-1 (src/mcx_host.cpp:759-768):
        # the exclusive prefix sum of the per-rank thread counts, which need not be equal (different devices / nthread)
        sim.reseed(p.c.seed, seed_offsets(dist, world, sim.nthread, device)[rank])
        sim.reset()
        sim.launch()
        detp, seeds, counts = combine(sim, dist, rank, world, device)
        if rank != 0:
            torch.cuda.synchronize()
            return None
        res = sim.fetch()
        if detp is not None:
            res["detp"] = _to_host(detp).reshape(-1, max(1, sim.reclen))
            res["detected"] = int(sum(counts))
            res["saved"] = res["detp"].shape[0]
            if seeds is not None:
                res["seeds"] = _to_host(seeds).view(np.uint64).reshape(-1, 2)
        res["nphoton"] = total
        res["shares"] = shares
        res["flux"] = engine.shape_field(p, res["field"])
        return res
