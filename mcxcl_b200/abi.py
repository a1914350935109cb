"""ctypes mirror of include/mcxb200.h and loader of the CUDA engine (libmcxb200.so).

There is deliberately no fallback: if the shared library was not built (``__graft_entry__.build()``
or ``python -m mcxcl_b200.build``) every entry point raises, it never reroutes to a CPU path.
"""
import ctypes as C
import os

ABI_VERSION = 4
NANGLES = 181
DEBUG_RNG = 1
DEBUG_MOVE, DEBUG_MOVE_ONLY = 2, 8
TRAJ_RECLEN = 6
DEBUG_STATS = 0x10000
ACCUM_F64, ACCUM_F32 = 0, 1
SCHED_DYNAMIC, SCHED_STATIC = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcxb200.so")


class F4(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]

    def __init__(self, x=0.0, y=0.0, z=0.0, w=0.0):
        super().__init__(float(x), float(y), float(z), float(w))

    def tolist(self):
        return [self.x, self.y, self.z, self.w]


class Source(C.Structure):
    _fields_ = [("pos", F4), ("dir", F4), ("param1", F4), ("param2", F4)]


class Config(C.Structure):
    """struct mcxb_config"""
    _fields_ = [
        ("abi_version", C.c_uint32),
        ("dimx", C.c_uint32), ("dimy", C.c_uint32), ("dimz", C.c_uint32),
        ("vol", C.POINTER(C.c_uint32)),
        ("unitinmm", C.c_float),
        ("mediaformat", C.c_uint32),
        ("medianum", C.c_uint32),
        ("prop", C.POINTER(F4)),
        ("srctype", C.c_int32),
        ("src", Source),
        ("extrasrclen", C.c_uint32),
        ("srcdata", C.POINTER(Source)),
        ("srcid", C.c_int32),
        ("srcnum", C.c_uint32),
        ("srcpattern", C.POINTER(C.c_float)),
        ("srcpattern_len", C.c_uint64),
        ("nphase", C.c_uint32), ("nangle", C.c_uint32),
        ("invcdf", C.POINTER(C.c_float)),
        ("angleinvcdf", C.POINTER(C.c_float)),
        ("detnum", C.c_uint32),
        ("detpos", C.POINTER(F4)),
        ("issavedet", C.c_int32),
        ("savedetflag", C.c_uint32),
        ("maxdetphoton", C.c_uint32),
        ("issaveseed", C.c_int32),
        ("issaveref", C.c_int32),
        ("tstart", C.c_float), ("tstep", C.c_float), ("tend", C.c_float),
        ("nphoton", C.c_uint64),
        ("seed", C.c_int32),
        ("seed_skip", C.c_uint64),
        ("isreflect", C.c_int32),
        ("bc", C.c_uint8 * 12),
        ("isspecular", C.c_int32),
        ("minenergy", C.c_float),
        ("gscatter", C.c_uint32),
        ("maxvoidstep", C.c_int32),
        ("voidtime", C.c_int32),
        ("outputtype", C.c_int32),
        ("isnormalized", C.c_int32),
        ("issave2pt", C.c_int32),
        ("debuglevel", C.c_uint32),
        ("nthread", C.c_uint32),
        ("nblocksize", C.c_uint32),
        ("sched", C.c_int32),
        ("accum", C.c_int32),
        ("replay_seed", C.POINTER(C.c_uint64)),
        ("replay_weight", C.POINTER(C.c_float)),
        ("replay_tof", C.POINTER(C.c_float)),
        ("replay_detid", C.POINTER(C.c_int32)),
        ("replaydet", C.c_int32),
        ("respin", C.c_int32),
        ("maxjumpdebug", C.c_uint32),
        ("polmedianum", C.c_uint32),
        ("smatrix", C.POINTER(F4)),
        ("srciquv", F4),
        ("omega", C.c_float),
    ]


class Output(C.Structure):
    """struct mcxb_output"""
    _fields_ = [
        ("field", C.POINTER(C.c_float)),
        ("fieldlen", C.c_uint64),
        ("detphoton", C.POINTER(C.c_float)),
        ("seeddata", C.POINTER(C.c_uint64)),
        ("detected", C.c_uint32),
        ("saved", C.c_uint32),
        ("reclen", C.c_uint32),
        ("maxgate", C.c_uint32),
        ("energytot", C.c_double), ("energyesc", C.c_double), ("energyabs", C.c_double),
        ("normalizer", C.c_float),
        ("runtime_ms", C.c_float),
        ("nthread", C.c_uint32), ("nblocksize", C.c_uint32),
        ("kernel_launches", C.c_uint64),
        ("stats", C.c_uint64 * 3),
        ("debugdata", C.POINTER(C.c_float)),
        ("debugrecorded", C.c_uint32),
        ("debugdatalen", C.c_uint32),
    ]


class GPUInfo(C.Structure):
    """struct mcxb_gpuinfo"""
    _fields_ = [
        ("name", C.c_char * 64),
        ("id", C.c_int32), ("devcount", C.c_int32), ("major", C.c_int32), ("minor", C.c_int32),
        ("globalmem", C.c_uint64), ("constmem", C.c_uint64), ("sharedmem", C.c_uint64),
        ("regcount", C.c_int32), ("clock_khz", C.c_int32), ("sm", C.c_int32), ("core", C.c_int32),
        ("autoblock", C.c_uint64), ("autothread", C.c_uint64),
        ("maxmpthread", C.c_int32),
        ("l2cache", C.c_uint64),
    ]


MAX_DEVICES = 16


class MultiInfo(C.Structure):
    """struct mcxb_multi_info"""
    _fields_ = [
        ("ndev", C.c_int32), ("nccl_version", C.c_int32),
        ("share", C.c_uint64 * MAX_DEVICES),
        ("detected", C.c_uint32 * MAX_DEVICES),
        ("nthread", C.c_uint32 * MAX_DEVICES),
        ("kernel_ms", C.c_float * MAX_DEVICES),
    ]


class TraceStep(C.Structure):
    """struct mcxb_trace_step"""
    _fields_ = [
        ("dist", C.c_float), ("px", C.c_float), ("py", C.c_float), ("pz", C.c_float),
        ("ix", C.c_int16), ("iy", C.c_int16), ("iz", C.c_int16), ("face", C.c_int16),
        ("idx1d", C.c_uint32),
    ]


# every symbol include/mcxb200.h declares: (name, restype, argtypes)
_VP = C.c_void_p
SYMBOLS = [
    ("mcxb_list_gpu", C.c_int, [C.POINTER(GPUInfo), C.c_int]),
    ("mcxb_run_simulation", C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(Output)]),
    ("mcxb_run_simulation_multi", C.c_int, [C.POINTER(Config), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_float), C.POINTER(Output), C.POINTER(MultiInfo)]),
    ("mcxb_split_photons", None, [C.c_uint64, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint64)]),
    ("mcxb_nccl_version", C.c_int, []),
    ("mcxb_last_error", C.c_char_p, []),
    ("mcxb_release_cached_buffers", None, []),
    ("mcxb_sim_create", C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(_VP)]),
    ("mcxb_sim_reset", C.c_int, [_VP, _VP]),
    ("mcxb_sim_launch", C.c_int, [_VP, _VP]),
    ("mcxb_sim_progress", C.c_int, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    ("mcxb_sim_set_photons", C.c_int, [_VP, C.c_uint64]),
    ("mcxb_sim_reseed", C.c_int, [_VP, C.c_int32, C.c_uint64]),
    ("mcxb_sim_run_batches", C.c_int, [_VP, C.c_uint64, C.c_uint32, C.c_int32, C.c_uint64, C.c_uint64, C.POINTER(C.c_float)]),
    ("mcxb_sim_finalize", C.c_int, [_VP, _VP]),
    ("mcxb_sim_fetch", C.c_int, [_VP, _VP, C.POINTER(Output)]),
    ("mcxb_sim_field_devptr", _VP, [_VP]),
    ("mcxb_sim_energy_devptr", _VP, [_VP]),
    ("mcxb_sim_detphoton_devptr", _VP, [_VP]),
    ("mcxb_sim_detcount_devptr", _VP, [_VP]),
    ("mcxb_sim_seeddata_devptr", _VP, [_VP]),
    ("mcxb_sim_fieldlen", C.c_uint64, [_VP]),
    ("mcxb_sim_reclen", C.c_uint32, [_VP]),
    ("mcxb_sim_nthread", C.c_uint32, [_VP]),
    ("mcxb_sim_acc_copies", C.c_uint32, [_VP]),
    ("mcxb_sim_kernel_name", C.c_char_p, [_VP]),
    ("mcxb_sim_last_kernel_ms", C.c_float, [_VP]),
    ("mcxb_sim_destroy", None, [_VP]),
    ("mcxb_normalizer", C.c_float, [C.POINTER(Config), C.c_double]),
    ("mcxb_adjoint_products", C.c_int, [C.c_int, _VP, _VP, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, _VP]),
    ("mcxb_test_rng", C.c_int, [C.c_int, _VP, C.c_uint32, C.c_uint32, _VP, _VP]),
    ("mcxb_test_trace", C.c_int, [C.c_int, _VP, _VP, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                  C.c_float, _VP]),
    ("mcxb_test_scalar", C.c_int, [C.c_int, _VP, _VP, C.c_uint32, _VP, _VP, _VP, _VP, _VP, C.c_uint32, _VP]),
    ("mcxb_test_rotate", C.c_int, [C.c_int, _VP, _VP, _VP, _VP, _VP, C.c_uint32]),
    ("mcxb_test_refract", C.c_int, [C.c_int, _VP, _VP, _VP, _VP, C.c_uint32]),
    ("mcxb_bench_red", C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                 C.POINTER(C.c_float), C.POINTER(C.c_uint64)]),
    ("mcxb_bench_red_die", C.c_int, [C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    ("mcxb_fill_seeds", None, [C.c_int32, C.c_uint64, C.c_uint64, _VP]),
]

_lib = None


class EngineMissing(RuntimeError):
    pass


def load():
    """Load libmcxb200.so (once) and bind every declared symbol; raises if the engine is not built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MCXB200_LIB", LIB_PATH)      # tuning experiments only (mcxcl_b200.build --variant)
    if not os.path.exists(path):
        raise EngineMissing(
            "CUDA engine %s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(code, what="mcxb200"):
    if code != 0:
        msg = load().mcxb_last_error()
        raise RuntimeError("%s failed with code %d: %s" % (what, code, (msg or b"").decode()))
