"""Shared helpers of the parity tests."""
import ctypes as C
import os

import numpy as np

from mcxcl_b200 import abi, benchmarks, engine, hostcfg

HERE = os.path.dirname(os.path.abspath(__file__))


def f32bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def gpu_trace(lib, p0, v0, nstep, dims, musp=1.0):
    from oracle.loader import TRACE_DTYPE
    p0 = np.ascontiguousarray(p0, dtype=np.float32).reshape(-1, 4)
    v0 = np.ascontiguousarray(v0, dtype=np.float32).reshape(-1, 4)
    n = p0.shape[0]
    buf = (abi.TraceStep * (n * nstep))()
    abi.check(lib.mcxb_test_trace(0, p0.ctypes.data, v0.ctypes.data, n, nstep, dims[0], dims[1], dims[2],
                                  C.c_float(musp), C.addressof(buf)), "mcxb_test_trace")
    return np.frombuffer(buf, dtype=TRACE_DTYPE).reshape(n, nstep).copy()


def records_equal(a, b):
    return all(a[f].tobytes() == b[f].tobytes() for f in a.dtype.names)


def random_rays(n, dims, seed):
    rs = np.random.RandomState(seed)
    p0 = np.zeros((n, 4), np.float32)
    p0[:, :3] = rs.uniform(0, 1, (n, 3)) * np.array(dims)
    v0 = np.zeros((n, 4), np.float32)
    d = rs.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    v0[:, :3] = d
    return p0, v0


def run_gpu(cfg, **over):
    c = dict(cfg)
    c.update(over)
    p = hostcfg.prepare(c)
    return p, engine.run_prepared(p)


def run_ref(checker, cfg, work=2048, **over):
    c = dict(cfg)
    c.update(over)
    p = hostcfg.prepare(c)
    return p, checker.run(p, work, hostthreads=0)


def absorbed_sigma(nphoton, absorbed, spread=1.6):
    """rough standard deviation of the absorbed fraction of an N-photon run; `spread` covers the packet-weight
    variance relative to a Bernoulli variable (calibrated on the committed reference series: cube60b at 2e5
    photons has sigma 7.3e-4, Bernoulli 1.0e-3)."""
    return spread * np.sqrt(max(absorbed * (1 - absorbed), 1e-4) / nphoton)


def zscores(field_raw, golden, scale=1.0):
    """z-scores of a raw (un-normalised) field against committed reference statistics."""
    idx = golden["idx"]
    runs = int(golden["runs"])
    mean, std = golden["mean"].astype(np.float64), golden["std"].astype(np.float64)
    x = field_raw.astype(np.float64)[idx] * scale
    ok = std > 0
    return (x[ok] - mean[ok]) / (std[ok] * np.sqrt(1.0 + 1.0 / runs)), ok


def bin_field(field, dims, b):
    nx, ny, nz = dims
    v = np.asarray(field, dtype=np.float64).reshape(nz, ny, nx)
    pz, py, px = (-nz) % b, (-ny) % b, (-nx) % b
    if pz or py or px:
        v = np.pad(v, ((0, pz), (0, py), (0, px)))
    return v.reshape((nz + pz) // b, b, (ny + py) // b, b, (nx + px) // b, b).sum(axis=(1, 3, 5)).ravel()
