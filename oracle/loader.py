"""ctypes loader for the two CPU checkers (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under mcxcl_b200/ does.
  ref()    -> oracle/_ref/libmcxref.so   the reference's own kernel source built for the host
  port()   -> oracle/libmcxoracle.so     the plain-C restatement (oracle/mcx_oracle.c)
"""
import ctypes as C
import os

import numpy as np

from mcxcl_b200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libmcxref.so")
PORT_LIB = os.path.join(HERE, "libmcxoracle.so")


class Result(C.Structure):
    """struct mcxo_result (oracle/oracle_api.h)"""
    _fields_ = [
        ("field", C.POINTER(C.c_float)), ("fieldlen", C.c_uint64),
        ("energy", C.POINTER(C.c_float)), ("detphoton", C.POINTER(C.c_float)),
        ("seeddata", C.POINTER(C.c_uint64)), ("detcap", C.c_uint32),
        ("detected", C.c_uint32), ("reclen", C.c_uint32),
        ("energytot", C.c_double), ("energyesc", C.c_double),
        ("n_segment", C.c_uint64), ("n_deposit", C.c_uint64), ("n_scatter", C.c_uint64), ("n_launch", C.c_uint64),
        ("runtime_ms", C.c_double),
        ("traj", C.POINTER(C.c_float)), ("trajcap", C.c_uint32), ("trajcount", C.c_uint32),
    ]


_VP = C.c_void_p
_libs = {}


def _load(path, prefix):
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise FileNotFoundError("%s not built (python oracle/build_ref.py / make -C oracle)" % path)
    lib = C.CDLL(path)
    sig = {
        "run": (C.c_int, [C.POINTER(abi.Config), C.c_uint32, C.c_int, C.POINTER(Result)]),
        "rng": (C.c_int, [_VP, C.c_uint32, C.c_uint32, _VP, _VP]),
        "trace": (C.c_int, [_VP, _VP, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_float, _VP]),
        "scalar": (C.c_int, [_VP, _VP, C.c_uint32, _VP, _VP, _VP, _VP, _VP, C.c_uint32, _VP]),
        "rotate": (C.c_int, [_VP, _VP, _VP, _VP, _VP, C.c_uint32]),
        "transmit": (C.c_int, [_VP, _VP, _VP, _VP, C.c_uint32]),
        "seeds": (None, [C.c_int, C.c_uint64, C.c_uint64, _VP]),
    }
    size = getattr(lib, prefix + "configsize")
    size.restype = C.c_uint
    if size() != C.sizeof(abi.Config):
        raise RuntimeError("%s was built against another include/mcxb200.h (sizeof(mcxb_config) %d != %d): rebuild it"
                           % (path, size(), C.sizeof(abi.Config)))
    for name, (res, args) in sig.items():
        fn = getattr(lib, prefix + name)
        fn.restype, fn.argtypes = res, args
    _libs[path] = lib
    return lib


class Checker:
    def __init__(self, lib, prefix, kind):
        self.lib, self.prefix, self.kind = lib, prefix, kind

    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def run(self, prepared, nthread, hostthreads=1, want_energy=False):
        """Run the CPU checker on a hostcfg.Prepared; returns a dict of numpy results."""
        p = prepared
        res = Result()
        field = np.zeros(p.fieldlen, dtype=np.float32)
        res.field = field.ctypes.data_as(C.POINTER(C.c_float))
        res.fieldlen = field.size
        energy = np.zeros(2 * nthread, dtype=np.float32) if want_energy else None
        if energy is not None:
            res.energy = energy.ctypes.data_as(C.POINTER(C.c_float))
        det = seeds = None
        if p.c.issavedet:
            det = np.zeros((p.c.maxdetphoton, max(1, p.reclen)), dtype=np.float32)
            res.detphoton = det.ctypes.data_as(C.POINTER(C.c_float))
            res.detcap = p.c.maxdetphoton
            if p.c.issaveseed:
                seeds = np.zeros((p.c.maxdetphoton, 2), dtype=np.uint64)
                res.seeddata = seeds.ctypes.data_as(C.POINTER(C.c_uint64))
        traj = None
        if p.c.debuglevel & 0xA:          # MCX_DEBUG_MOVE / MCX_DEBUG_MOVE_ONLY
            traj = np.zeros((p.c.maxjumpdebug, 6), dtype=np.float32)
            res.traj = traj.ctypes.data_as(C.POINTER(C.c_float))
            res.trajcap = p.c.maxjumpdebug
        rc = self._fn("run")(C.byref(p.c), int(nthread), int(hostthreads), C.byref(res))
        if rc != 0:
            raise RuntimeError("%s run failed: %d" % (self.kind, rc))
        nsaved = min(res.detected, p.c.maxdetphoton) if det is not None else 0
        return dict(field=field, energy=energy, energytot=res.energytot, energyesc=res.energyesc,
                    absorbed=(res.energytot - res.energyesc) / res.energytot if res.energytot else 0.0,
                    detected=res.detected, detp=None if det is None else det[:nsaved],
                    seeds=None if seeds is None else seeds[:nsaved], reclen=res.reclen,
                    n_segment=res.n_segment, n_deposit=res.n_deposit, n_scatter=res.n_scatter,
                    n_launch=res.n_launch, runtime_ms=res.runtime_ms, traj=None if traj is None else traj[:res.trajcount])

    def seeds(self, seed, nrecords, skip=0):
        out = np.zeros(4 * nrecords, dtype=np.uint32)
        self._fn("seeds")(int(seed), int(skip), int(nrecords), out.ctypes.data)
        return out

    def rng(self, seeds, ndraw):
        seeds = np.ascontiguousarray(seeds, dtype=np.uint32)
        n = seeds.size // 4
        out = np.zeros((n, ndraw), dtype=np.float32)
        state = np.zeros((n, 2), dtype=np.uint64)
        self._fn("rng")(seeds.ctypes.data, n, ndraw, out.ctypes.data, state.ctypes.data)
        return out, state

    def trace(self, p0, v0, nstep, dims, musp=1.0):
        p0 = np.ascontiguousarray(p0, dtype=np.float32).reshape(-1, 4)
        v0 = np.ascontiguousarray(v0, dtype=np.float32).reshape(-1, 4)
        n = p0.shape[0]
        out = (abi.TraceStep * (n * nstep))()
        self._fn("trace")(p0.ctypes.data, v0.ctypes.data, n, nstep, dims[0], dims[1], dims[2], float(musp), C.addressof(out))
        return np.frombuffer(out, dtype=TRACE_DTYPE).reshape(n, nstep).copy()

    def scalar(self, a, direction, v, n1, n2, face):
        a = np.ascontiguousarray(a, dtype=np.float32)
        direction = np.ascontiguousarray(direction, dtype=np.int32)
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 4)
        n1 = np.ascontiguousarray(n1, dtype=np.float32)
        n2 = np.ascontiguousarray(n2, dtype=np.float32)
        face = np.ascontiguousarray(face, dtype=np.int32)
        na = np.zeros(a.size, dtype=np.float32)
        rc = np.zeros(v.shape[0], dtype=np.float32)
        self._fn("scalar")(a.ctypes.data, direction.ctypes.data, a.size, na.ctypes.data,
                           v.ctypes.data, n1.ctypes.data, n2.ctypes.data, face.ctypes.data, v.shape[0], rc.ctypes.data)
        return na, rc

    def rotate(self, v, st, ct, sp, cp):
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 4).copy()
        arrs = [np.ascontiguousarray(x, dtype=np.float32) for x in (st, ct, sp, cp)]
        self._fn("rotate")(v.ctypes.data, *[x.ctypes.data for x in arrs], v.shape[0])
        return v

    def transmit(self, v, n1, n2, face):
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 4).copy()
        n1 = np.ascontiguousarray(n1, dtype=np.float32)
        n2 = np.ascontiguousarray(n2, dtype=np.float32)
        face = np.ascontiguousarray(face, dtype=np.int32)
        self._fn("transmit")(v.ctypes.data, n1.ctypes.data, n2.ctypes.data, face.ctypes.data, v.shape[0])
        return v


TRACE_DTYPE = np.dtype([("dist", "<f4"), ("px", "<f4"), ("py", "<f4"), ("pz", "<f4"),
                        ("ix", "<i2"), ("iy", "<i2"), ("iz", "<i2"), ("face", "<i2"), ("idx1d", "<u4")])


def ref():
    return Checker(_load(REF_LIB, "mcxref_"), "mcxref_", "reference")


def port():
    return Checker(_load(PORT_LIB, "mcxo_"), "mcxo_", "port")


def have_ref():
    return os.path.exists(REF_LIB)


def have_port():
    return os.path.exists(PORT_LIB)
