"""Host-side driver of the CUDA engine: the Python face of the reference's ``pmcxcl.run`` / ``mcx_run_simulation``.

``run(cfg)`` takes the same dictionary as ``pmcxcl.run(**cfg)`` (reference src/pmcxcl.cpp:1124-1555) and
returns the same result dictionary -- ``flux`` (Nx,Ny,Nz,Ngate[,Nsrc]) float32, ``detp`` (reclen, ndet),
``seeds``, ``stat`` {runtime, nphoton, energytot, energyabs, normalizer, unitinmm, workload} (:1494-1509)
-- computed by libmcxb200.so through the C ABI of include/mcxb200.h.  ``Simulation`` is the staged form
(inputs resident in HBM) used by bench.py and by the multi-GPU driver in mcxcl_b200.multigpu.

There is no CPU path here: abi.load() raises if the CUDA library has not been built.
"""
import ctypes as C

import numpy as np

from . import abi, hostcfg


def gpuinfo():
    """List CUDA devices like ``pmcxcl.gpuinfo()`` (reference src/pmcxcl.cpp:1600-1646)."""
    lib = abi.load()
    buf = (abi.GPUInfo * 16)()
    n = lib.mcxb_list_gpu(buf, 16)
    if n < 0:
        abi.check(n, "mcxb_list_gpu")
    out = []
    for i in range(min(n, 16)):
        g = buf[i]
        out.append(dict(name=g.name.decode(), id=g.id, devcount=g.devcount, major=g.major, minor=g.minor,
                        globalmem=g.globalmem, constmem=g.constmem, sharedmem=g.sharedmem, regcount=g.regcount,
                        clock=g.clock_khz, sm=g.sm, core=g.core, autoblock=g.autoblock, autothread=g.autothread,
                        maxgate=0, l2cache=g.l2cache))
    return out


class Simulation:
    """One prepared simulation resident on one GPU (mcxb_sim_* of include/mcxb200.h)."""

    def __init__(self, prepared, device=0):
        self.lib = abi.load()
        self.p = prepared
        self.device = int(device)
        self.h = C.c_void_p()
        abi.check(self.lib.mcxb_sim_create(C.byref(prepared.c), self.device, C.byref(self.h)), "mcxb_sim_create")
        self.fieldlen = int(self.lib.mcxb_sim_fieldlen(self.h))
        self.reclen = int(self.lib.mcxb_sim_reclen(self.h))
        self.nthread = int(self.lib.mcxb_sim_nthread(self.h))
        self.acc_copies = int(self.lib.mcxb_sim_acc_copies(self.h))
        self.kernel_name = self.lib.mcxb_sim_kernel_name(self.h).decode()

    def close(self):
        if self.h:
            self.lib.mcxb_sim_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def progress(self):
        """(photons claimed so far, launch finished) without waiting for the kernel -- the `-D P` poll"""
        n, done = C.c_uint64(0), C.c_int(0)
        abi.check(self.lib.mcxb_sim_progress(self.h, C.byref(n), C.byref(done)), "mcxb_sim_progress")
        return int(n.value), bool(done.value)

    def set_photons(self, n):
        abi.check(self.lib.mcxb_sim_set_photons(self.h, int(n)), "mcxb_sim_set_photons")

    def reseed(self, seed, seed_skip=0):
        abi.check(self.lib.mcxb_sim_reseed(self.h, int(seed), int(seed_skip)), "mcxb_sim_reseed")

    def run_batches(self, nphoton, respin, seed, seed_skip=0, seed_stride=None):
        """Config.respin on this resident simulation (mcxb_sim_run_batches); returns the summed kernel time in ms"""
        ms = C.c_float(0)
        abi.check(self.lib.mcxb_sim_run_batches(self.h, int(nphoton), int(respin), int(seed), int(seed_skip),
                                                int(self.nthread if seed_stride is None else seed_stride), C.byref(ms)), "mcxb_sim_run_batches")
        return float(ms.value)

    def reset(self, stream=None):
        abi.check(self.lib.mcxb_sim_reset(self.h, C.c_void_p(stream or 0)), "mcxb_sim_reset")

    def launch(self, stream=None):
        abi.check(self.lib.mcxb_sim_launch(self.h, C.c_void_p(stream or 0)), "mcxb_sim_launch")

    def finalize(self, stream=None):
        abi.check(self.lib.mcxb_sim_finalize(self.h, C.c_void_p(stream or 0)), "mcxb_sim_finalize")

    def kernel_ms(self):
        return float(self.lib.mcxb_sim_last_kernel_ms(self.h))

    def fetch(self, stream=None, field=None, want_field=True):
        """Sync, read back and normalise.  `field` (float32, fieldlen) is accumulated into, like cfg->exportfield."""
        out = abi.Output()
        if want_field:
            if field is None:
                field = np.zeros(self.fieldlen, dtype=np.float32)
            assert field.dtype == np.float32 and field.size >= self.fieldlen and field.flags.c_contiguous
            out.field = field.ctypes.data_as(C.POINTER(C.c_float))
            out.fieldlen = field.size
        det = seeds = None
        c = self.p.c
        if c.issavedet and not (c.debuglevel & 1):
            det = np.zeros((c.maxdetphoton, max(1, self.reclen)), dtype=np.float32)
            out.detphoton = det.ctypes.data_as(C.POINTER(C.c_float))
            if c.issaveseed:
                seeds = np.zeros((c.maxdetphoton, 2), dtype=np.uint64)
                out.seeddata = seeds.ctypes.data_as(C.POINTER(C.c_uint64))
        traj = _traj_buffer(c, out)
        abi.check(self.lib.mcxb_sim_fetch(self.h, C.c_void_p(stream or 0), C.byref(out)), "mcxb_sim_fetch")
        return _result(self.p, out, field if want_field else None, det, seeds, traj)

    # raw device pointers for the multi-GPU reducer
    def devptrs(self):
        L, h = self.lib, self.h
        return dict(field=L.mcxb_sim_field_devptr(h), energy=L.mcxb_sim_energy_devptr(h),
                    detphoton=L.mcxb_sim_detphoton_devptr(h), detcount=L.mcxb_sim_detcount_devptr(h),
                    seeddata=L.mcxb_sim_seeddata_devptr(h))


def _traj_buffer(c, out):
    """caller-owned trajectory buffer when `-D M` / `-D T` is on (cfg->exportdebugdata, src/pmcxcl.cpp:1229-1231)"""
    if not (c.debuglevel & (abi.DEBUG_MOVE | abi.DEBUG_MOVE_ONLY)) or (c.debuglevel & abi.DEBUG_RNG):
        return None
    traj = np.zeros((max(1, c.maxjumpdebug), abi.TRAJ_RECLEN), dtype=np.float32)
    out.debugdata = traj.ctypes.data_as(C.POINTER(C.c_float))
    return traj


def _result(p, out, field, det, seeds, traj=None):
    c = p.c
    res = dict(
        energytot=out.energytot, energyesc=out.energyesc, energyabs=out.energyabs,
        absorbed=(out.energyabs / out.energytot) if out.energytot else 0.0,
        detected=int(out.detected), saved=int(out.saved), reclen=int(out.reclen), maxgate=int(out.maxgate),
        normalizer=float(out.normalizer), runtime_ms=float(out.runtime_ms), nthread=int(out.nthread),
        nblocksize=int(out.nblocksize), kernel_launches=int(out.kernel_launches),
        stats=dict(segments=int(out.stats[0]), deposits=int(out.stats[1]), scatters=int(out.stats[2])),
        field=field, detp=None, seeds=None)
    if det is not None:
        res["detp"] = det[:out.saved]
        if seeds is not None:
            res["seeds"] = seeds[:out.saved]
    if traj is not None:
        res["traj"] = traj[:out.debugdatalen]          # rows: {photon id (uint32 bits), x, y, z, weight, source id}
        res["traj_recorded"] = int(out.debugrecorded)
    return res


def run_prepared(prepared, device=0, field=None):
    """One-shot call through mcxb_run_simulation with HOST buffers (the e2e path)."""
    lib = abi.load()
    p = prepared
    c = p.c
    out = abi.Output()
    if field is None:
        field = np.zeros(p.fieldlen, dtype=np.float32)
    out.field = field.ctypes.data_as(C.POINTER(C.c_float))
    out.fieldlen = field.size
    det = seeds = None
    if c.issavedet and not (c.debuglevel & 1):
        det = np.zeros((c.maxdetphoton, max(1, p.reclen)), dtype=np.float32)
        out.detphoton = det.ctypes.data_as(C.POINTER(C.c_float))
        if c.issaveseed:
            seeds = np.zeros((c.maxdetphoton, 2), dtype=np.uint64)
            out.seeddata = seeds.ctypes.data_as(C.POINTER(C.c_uint64))
    traj = _traj_buffer(c, out)
    abi.check(lib.mcxb_run_simulation(C.byref(c), int(device), C.byref(out)), "mcxb_run_simulation")
    return _result(p, out, field, det, seeds, traj)


def run_prepared_multi(prepared, devices, workload=None, field=None):
    """One call, several GPUs of this box, one host process: mcxb_run_simulation_multi (photon shards, NCCL reduce of
    the volume and energies, gather of the detected-photon records onto devices[0]).  Returns the result dictionary of
    run_prepared plus `multi` = the per-device report."""
    lib = abi.load()
    p = prepared
    c = p.c
    out = abi.Output()
    if field is None:
        field = np.zeros(p.fieldlen, dtype=np.float32)
    out.field = field.ctypes.data_as(C.POINTER(C.c_float))
    out.fieldlen = field.size
    det = seeds = None
    if c.issavedet and not (c.debuglevel & 1):
        det = np.zeros((c.maxdetphoton, max(1, p.reclen)), dtype=np.float32)
        out.detphoton = det.ctypes.data_as(C.POINTER(C.c_float))
        if c.issaveseed:
            seeds = np.zeros((c.maxdetphoton, 2), dtype=np.uint64)
            out.seeddata = seeds.ctypes.data_as(C.POINTER(C.c_uint64))
    devs = (C.c_int * len(devices))(*[int(d) for d in devices])
    wl = None
    if workload is not None:
        assert len(workload) == len(devices)
        wl = (C.c_float * len(devices))(*[float(w) for w in workload])
    info = abi.MultiInfo()
    abi.check(lib.mcxb_run_simulation_multi(C.byref(c), devs, len(devices), wl, C.byref(out), C.byref(info)), "mcxb_run_simulation_multi")
    res = _result(p, out, field, det, seeds)
    n = info.ndev
    res["multi"] = dict(ndev=n, nccl_version=info.nccl_version, share=[int(info.share[i]) for i in range(n)],
                        detected=[int(info.detected[i]) for i in range(n)], nthread=[int(info.nthread[i]) for i in range(n)],
                        kernel_ms=[float(info.kernel_ms[i]) for i in range(n)])
    return res


def shape_field(p, field):
    """float32[fieldlen] -> (Nx,Ny,Nz,Ngate[,Nsrc]) view, the layout pmcxcl returns (column-major file order
    [Nx][Ny][Nz][Ng][Ns], README.md:1316-1323)."""
    nx, ny, nz = p.dims
    if p.c.srcnum > 1:          # photon sharing: patterns interleaved fastest, pmcxcl returns (srcnum*Nx, Ny, Nz, Ng) (src/pmcxcl.cpp:1330)
        return field[:p.fieldlen].reshape((p.c.srcnum * nx, ny, nz, p.maxgate), order="F")
    shp = (nx, ny, nz, p.maxgate) + ((p.nrepvol,) if p.nrepvol > 1 else ()) + ((p.nsrcvol,) if p.nsrcvol > 1 else ())
    if p.rfplanes == 2:         # RF outputs: real volumes, then imaginary volumes -> one complex array
        half = p.fieldlen // 2
        return (field[:half] + 1j * field[half:p.fieldlen]).astype(np.complex64).reshape(shp, order="F")
    return field[:p.fieldlen].reshape(shp, order="F")


def run(cfg=None, device=0, **kw):
    """Drop-in for ``pmcxcl.run(cfg)`` / ``pmcxcl.run(**cfg)`` on one B200."""
    cfg = dict(cfg or {})
    cfg.update(kw)
    gpuid = cfg.pop("gpuid", None)
    if isinstance(gpuid, int) and gpuid > 0:
        device = gpuid - 1
    p = hostcfg.prepare(cfg)
    r = run_prepared(p, device)
    out = dict(flux=shape_field(p, r["field"]))
    if r["detp"] is not None:
        out["detp"] = np.ascontiguousarray(r["detp"].T)
        if r["seeds"] is not None:
            out["seeds"] = r["seeds"].view(np.uint8).reshape(-1, 16).T.copy()
    if r.get("traj") is not None:
        out["traj"] = np.asfortranarray(r["traj"].T)            # (6, N) like pmcxcl (src/pmcxcl.cpp:1266-1282)
    out["stat"] = dict(runtime=r["runtime_ms"], nphoton=int(p.c.nphoton), energytot=r["energytot"],
                       energyabs=r["energyabs"], normalizer=r["normalizer"], unitinmm=p.c.unitinmm,
                       workload=[1.0], detected=r["detected"], absorbed=r["absorbed"])
    return out
