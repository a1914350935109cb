"""Generate tests/golden/*.npz from the reference's own kernel source (oracle/_ref/libmcxref.so).

    python tests/golden/make_golden.py

Needs /root/reference (to build oracle/_ref); the outputs are committed so the GPU box, which has no
reference tree, can still test against reference statistics.

ref_stats_<deck>.npz:  per-voxel mean and seed-to-seed standard deviation of the RAW (un-normalised)
fluence deposits of R independent reference runs (seeds s0 .. s0+R-1, N photons, W work-items), kept for
the voxels whose mean exceeds 1e-4 of the peak (the criterion of BASELINE.json), plus the absorbed
fraction, detected count and per-photon work counters of every run.

The three large decks (skinvessel, colin27, digimouse: 7-10 M voxels) are run at 1e6 photons x 8 seeds and keep
(a) `idx/mean/std`: VOXEL-wise statistics of the voxels above 1e-4 of the peak -- when more than MAXVOX voxels
qualify, a deterministic hash subsample of them (`subsample` = the modulus; the test draws nothing itself), and
(b) `bidx/bmean/bstd`: the same statistics on 8x8x8-voxel blocks, which cover the whole volume.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mcxcl_b200 import benchmarks, hostcfg  # noqa: E402
from oracle import loader  # noqa: E402

DECKS = {
    # name: (deck kwargs, photons per run, work-items, runs)
    "cube60": (dict(name="cube60"), 200000, 4096, 12),
    "cube60b": (dict(name="cube60b"), 200000, 4096, 12),
    # 200^3 volume and the two atlases (no pins in the reference's own tests for the atlases: the reference series
    # IS the pin); voxel-wise + 8x8x8 blocks (dimensions zero-padded up to a multiple of the block size)
    "skinvessel": (dict(name="skinvessel", bin=8), 1000000, 4096, 8),
    "colin27": (dict(name="colin27", bin=8), 1000000, 4096, 8),
    "digimouse": (dict(name="digimouse", bin=8), 1000000, 4096, 8),
}
SEED0 = 1648335518
MAXVOX = 200000


def voxel_hash(idx):
    """deterministic 32-bit mix of the voxel index (Knuth multiplicative), used to subsample large voxel sets"""
    return ((idx.astype(np.uint64) * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)) >> np.uint64(7)


def bin_field(field, dims, b):
    """sum an x-fastest flat volume over b x b x b blocks"""
    nx, ny, nz = dims
    v = field.reshape(nz, ny, nx)
    pz, py, px = (-nz) % b, (-ny) % b, (-nx) % b
    if pz or py or px:
        v = np.pad(v, ((0, pz), (0, py), (0, px)))
    return v.reshape((nz + pz) // b, b, (ny + py) // b, b, (nx + px) // b, b).sum(axis=(1, 3, 5)).ravel()


def main():
    ref = loader.ref()
    only = sys.argv[1:]
    for key, (kw, nph, work, runs) in DECKS.items():
        if only and key not in only:
            continue
        cfg = benchmarks.get(kw["name"], nph)
        fields, absorbed, detected, seg, dep, sca, etot = [], [], [], [], [], [], []
        big = bool(kw.get("bin"))
        s1 = s2 = None                      # running sums of the voxel-wise field (large decks: 8 runs x 10 M voxels)
        totals = []
        for r in range(runs):
            cfg["seed"] = SEED0 + r
            p = hostcfg.prepare(cfg)
            o = ref.run(p, work, hostthreads=0)
            fld = o["field"].astype(np.float64)
            totals.append(fld.sum())
            if big:
                s1 = fld.copy() if s1 is None else s1 + fld
                s2 = fld * fld if s2 is None else s2 + fld * fld
                fld = bin_field(fld, p.dims, kw["bin"])
            fields.append(fld)
            absorbed.append(o["absorbed"])
            detected.append(o["detected"])
            etot.append(o["energytot"])
            seg.append(o["n_segment"] / nph)
            dep.append(o["n_deposit"] / nph)
            sca.append(o["n_scatter"] / nph)
            print(key, r, o["absorbed"], o["detected"], flush=True)
        f = np.stack(fields)
        mean, std = f.mean(0), f.std(0, ddof=1)
        keep = np.nonzero(mean > 1e-4 * mean.max())[0]
        extra = {}
        if big:
            extra = dict(bidx=keep.astype(np.uint32), bmean=mean[keep].astype(np.float32), bstd=std[keep].astype(np.float32))
            mean = s1 / runs
            std = np.sqrt(np.maximum(s2 - runs * mean * mean, 0.0) / (runs - 1))
            keep = np.nonzero(mean > 1e-4 * mean.max())[0]
            extra["nabove"] = keep.size
            m = max(1, -(-keep.size // MAXVOX))
            keep = keep[voxel_hash(keep) % np.uint64(m) == 0]
            extra["subsample"] = m
        np.savez_compressed(os.path.join(HERE, "ref_stats_%s.npz" % key),
                            idx=keep.astype(np.uint32), mean=mean[keep].astype(np.float32), std=std[keep].astype(np.float32),
                            total=np.array(totals), absorbed=np.array(absorbed), detected=np.array(detected),
                            seg=np.array(seg), dep=np.array(dep), sca=np.array(sca), energytot=np.array(etot),
                            nphoton=nph, work=work, runs=runs, seed0=SEED0, bin=kw.get("bin", 1), **extra)
        print(key, "voxels kept", keep.size, "absorbed", np.mean(absorbed), "+-", np.std(absorbed, ddof=1))


if __name__ == "__main__":
    main()
