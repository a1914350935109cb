"""Host-side mirror of the reference's configuration layer for the photon-transport path.

`prepare(cfg_dict)` accepts the same keys as ``pmcxcl.run(**cfg)`` (reference src/pmcxcl.cpp:428-1110),
applies the defaults of ``mcx_initcfg`` (src/mcx_utils.c:203-366), the checks of ``mcx_validatecfg``
(:1822-1964) and the transformations of ``mcx_preprocess`` (:1521-1809) and ``mcx_maskdet``
(:4085-4198), and returns a `Prepared` object that owns the numpy buffers and the ctypes
``mcxb_config`` handed to the C ABI (include/mcxb200.h).  Only what the hot path consumes is
mirrored (including the replay set-up of ``mcx_replayinit``, :1355-1421); file formats, shapes and polarised input stay with the reference's own host code
(INTEGRATION.md shows how that code binds to the same ABI).
"""
import ctypes as C
import math

import numpy as np

from . import abi

SRCTYPES = ["pencil", "isotropic", "cone", "gaussian", "planar", "pattern", "fourier", "arcsine", "disk",
            "fourierx", "fourierx2d", "zgaussian", "line", "slit", "pencilarray", "pattern3d", "hyperboloid", "ring"]
# names of src/pmcxcl.cpp:854-872 (wl -> jacobian, wp -> nscat, wm = momentum transfer) and the -O letters of src/mcx_utils.c:143
OUTPUTTYPES = {"flux": 0, "fluence": 1, "energy": 2, "jacobian": 3, "wl": 3, "nscat": 4, "wp": 4, "wm": 5, "rf": 6, "length": 7, "rfmus": 8,
               "wltof": 9, "wptof": 10, "adjoint": 11, "adjoint_dcoeff": 12, "adjoint_mus": 13, "adjoint_musp": 14, "adjoint_mua_d": 15,
               "adjoint_mua_musp": 16,
               "x": 0, "f": 1, "e": 2, "j": 3, "p": 4, "m": 5, "r": 6, "l": 7, "s": 8, "t": 9, "b": 10}
REPLAY_OUTPUTS = (3, 4, 5, 6, 8, 9, 10)
SEED_FROM_FILE = -999        # src/mcx_const.h
R_C0 = np.float32(3.335640951981520e-12)   # 1/c0 in s/mm
BC_CODES = "_ramc"           # boundarycond[] (src/mcx_utils.c:170)
SAVEFLAGS = "DSPMXVWI"       # saveflag[]     (src/mcx_utils.c:134)
DET_MASK = 0x80000000
MED_MASK = 0x7FFFFFFF
HOST_EPS = np.float32(1e-10)  # EPS of src/mcx_const.h:33 (host side)


class ConfigError(ValueError):
    """Raised where the reference calls MCX_ERROR(id, msg); .code carries the id."""

    def __init__(self, code, msg):
        super().__init__("MCXCL ERROR(%d):%s" % (code, msg))
        self.code = code


def _f4(v, w_default=0.0):
    v = [float(x) for x in np.asarray(v, dtype=np.float64).ravel()]
    if len(v) == 3:
        v.append(w_default)
    if len(v) != 4:
        raise ConfigError(-6, "vector fields must have 3 or 4 elements")
    return np.array(v, dtype=np.float32)


def parse_savedetflag(flag):
    if isinstance(flag, str):
        out = 0
        for ch in flag.upper():
            i = SAVEFLAGS.find(ch)
            if i < 0:
                raise ConfigError(-6, "unknown savedetflag letter %r" % ch)
            out |= 1 << i
        return out
    return int(flag)


def parse_bc(bc):
    """'aarraa' / '______111111' -> 12 integer codes (mcx_lookupindex, src/mcx_utils.c:1562-1575)."""
    codes = np.zeros(12, dtype=np.uint8)
    if bc is None:
        return codes
    if not isinstance(bc, str):
        arr = np.asarray(bc, dtype=np.uint8).ravel()
        codes[:len(arr)] = arr[:12]
        return codes
    for i, ch in enumerate(bc[:12]):
        if i < 6:
            k = BC_CODES.find(ch)
            if k < 0:
                raise ConfigError(-4, "unknown boundary condition specifier")
        else:
            k = "01".find(ch)
            if k < 0:
                raise ConfigError(-4, "unknown boundary detection flags")
        codes[i] = k
    return codes


def maskdet(vol, dims, detpos, svmc=False):
    """Flag the surface voxels covered by each detector sphere in bit 31 (mcx_maskdet).

    vol: uint32[dimxyz] x-fastest (modified in place); returns the per-detector voxel counts.
    Follows src/mcx_utils.c:4085-4198 statement by statement, in float32 like the C code.
    """
    nx, ny, nz = dims
    f32 = np.float32
    pad = np.zeros((nz + 2, ny + 2, nx + 2), dtype=np.uint32)
    pad[1:-1, 1:-1, 1:-1] = vol.reshape(nz, ny, nx)
    isonecube = (nx == 1 and ny == 1 and nz == 1)
    corners = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1), (1, 1, 1)]
    nb = [(dz, dy, dx) for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1) if (dz, dy, dx) != (0, 0, 0)]
    counts = []
    vol3 = vol.reshape(nz, ny, nx)
    for d in range(len(detpos)):
        cx, cy, cz, r = [f32(t) for t in detpos[d]]
        count = 0
        d2max = (r + f32(1.7321)) * (r + f32(1.7321))
        lim = (r + f32(0.5) * f32(1 + isonecube)) ** 2
        z = -r - f32(1)
        while z <= r + f32(1):
            iz = z + cz
            y = -r - f32(1)
            while y <= r + f32(1):
                iy = y + cy
                x = -r - f32(1)
                while x <= r + f32(1):
                    ix = x + cx
                    xx = x
                    x = x + f32(0.5)
                    if (iz < 0 or ix < 0 or iy < 0 or ix >= nx or iy >= ny or iz >= nz or
                            xx * xx + y * y + z * z > (r + f32(1)) * (r + f32(1))):
                        continue
                    mind2 = None
                    for c in corners:
                        rx = f32(int(ix)) - cx + f32(c[0])
                        ry = f32(int(iy)) - cy + f32(c[1])
                        rz = f32(int(iz)) - cz + f32(c[2])
                        d2 = rx * rx + ry * ry + rz * rz
                        if d2 > d2max:
                            mind2 = None
                            break
                        if mind2 is None or d2 < mind2:
                            mind2 = d2
                    if mind2 is None or mind2 >= lim:
                        continue
                    pz, py, px = int(iz + f32(1)), int(iy + f32(1)), int(ix + f32(1))
                    if svmc:
                        # split-voxel media (src/mcx_utils.c:4157-4168): a split voxel whose lower part is empty
                        w = int(pad[pz, py, px])
                        if not (w >> 24) & 0x7F and (w >> 16) & 0xFF:
                            vol3[int(iz), int(iy), int(ix)] |= np.uint32(DET_MASK)
                            count += 1
                    elif pad[pz, py, px]:
                        if not all(pad[pz + a, py + b, px + c2] for a, b, c2 in nb):
                            vol3[int(iz), int(iy), int(ix)] |= np.uint32(DET_MASK)
                            count += 1
                y = y + f32(0.5)
            z = z + f32(0.5)
        counts.append(count)
    return counts


class Prepared:
    """Validated, pre-processed simulation input; owns every buffer the ctypes config points to."""

    def __init__(self):
        self.c = abi.Config()
        self.keep = {}
        self.session = ""
        self.det_voxels = []

    # derived quantities, computed exactly like src/mcx_host.cpp:474, 494-496, 647
    @property
    def dims(self):
        return (self.c.dimx, self.c.dimy, self.c.dimz)

    @property
    def dimxyz(self):
        return self.c.dimx * self.c.dimy * self.c.dimz

    @property
    def maxgate(self):
        return int((np.float32(self.c.tend) - np.float32(self.c.tstart)) / np.float32(self.c.tstep) + 0.5)

    @property
    def nsrcvol(self):
        # same rule as the engine (csrc/engine.cu, src/mcx_host.cpp:679): one volume per pattern with photon sharing,
        # else one per source when srcid < 0
        if self.c.srcnum > 1:
            return self.c.srcnum
        return self.c.extrasrclen + 1 if self.c.srcid < 0 else 1

    @property
    def nrepvol(self):
        return max(1, self.c.detnum) if (bool(self.c.replay_seed) and self.c.replaydet == -1) else 1

    @property
    def rfplanes(self):
        # RF outputs are complex: real volumes, then imaginary volumes (src/pmcxcl.cpp:1209-1211)
        replay = bool(self.c.replay_seed)
        return 2 if ((self.c.omega > 0 and not replay) or (replay and self.c.outputtype in (6, 8))) else 1

    @property
    def fieldlen(self):
        return self.dimxyz * self.maxgate * self.nsrcvol * self.nrepvol * self.rfplanes

    @property
    def partialdata(self):
        f = self.c.savedetflag if self.c.issavedet else 0
        return (self.c.medianum - 1) * ((f >> 1 & 1) + (f >> 2 & 1) + (f >> 3 & 1))

    @property
    def reclen(self):
        f = self.c.savedetflag if self.c.issavedet else 0
        return self.partialdata + (f & 1) + 3 * ((f >> 4 & 1) + (f >> 5 & 1)) + (f >> 6 & 1) + 4 * (f >> 7 & 1)

    def clone_for(self, nphoton=None, seed=None, seed_skip=None, **overrides):
        """Shallow copy sharing the big buffers, with a different photon budget / seed slice."""
        other = Prepared()
        C.memmove(C.byref(other.c), C.byref(self.c), C.sizeof(abi.Config))
        other.keep = self.keep
        other.session = self.session
        other.det_voxels = self.det_voxels
        if nphoton is not None:
            other.c.nphoton = int(nphoton)
        if seed is not None:
            other.c.seed = int(seed)
        if seed_skip is not None:
            other.c.seed_skip = int(seed_skip)
        for k, v in overrides.items():
            setattr(other.c, k, v)
        return other


MEDIA_LABEL_HALF, MEDIA_AS_F2H, MEDIA_MUA_FLOAT, MEDIA_AS_HALF, MEDIA_ASGN_BYTE, MEDIA_AS_SHORT = 99, 100, 101, 102, 103, 104
MEDIA_2LABEL_SPLIT = 97


def _float_to_half_bits(f):
    """float32 -> IEEE half bit patterns the way pmcxcl packs continuous media (src/pmcxcl.cpp:286-318): mantissa
    TRUNCATED to 10 bits, exponent re-biased and clamped, denormals (2^-24 .. 2^-14) rounded from an 11-bit mantissa"""
    i = np.ascontiguousarray(f, dtype=np.float32).view(np.uint32).astype(np.int64)
    m = (i >> 13) & 0x3FF
    e = (i >> 23) & 0xFF
    tmp = np.where(e > 0x70, (e - 0x70) & 0x1F, 0)
    normal = (((i >> 31) << 5 | tmp) << 10) | m
    sign = (i >> 16) & 0x8000
    m11 = ((i >> 12) & 0x7FF) | 0x800
    sh = (114 - e) & 31                         # the C code shifts an int by this count: taken mod 32 on x86
    den = sign | ((m11 >> sh) + ((m11 >> ((113 - e) & 31)) & 1))
    return np.where((m < 0x10) & (tmp == 0), den, normal).astype(np.uint32) & 0xFFFF


def pack_continuous_volume(vol, unitinmm=1.0):
    """4-D `vol` of pmcxcl (component, x, y, z) -> (uint32 words shaped (x, y, z), media format), mirroring the packing of
    src/pmcxcl.cpp:108-400: float32 x1 = mua (MEDIA_MUA_FLOAT), float32 x2 = mua, mus as halves (MEDIA_AS_F2H), float32 x3 =
    {value, slot 0..3, label} (MEDIA_LABEL_HALF), int8/uint8 x4 = mua, mus, g, n bytes (MEDIA_ASGN_BYTE), int16/uint16 x2 =
    mua, mus shorts (MEDIA_AS_SHORT).  mua and mus are scaled by unitinmm where pmcxcl does."""
    v = np.asarray(vol)
    if v.ndim != 4:
        raise ConfigError(-4, "a continuous-media volume is 4-D: (component, x, y, z)")
    ch = v.shape[0]
    u = np.float32(unitinmm)
    if v.dtype in (np.int8, np.uint8) and ch == 8:
        # split-voxel media (src/pmcxcl.cpp:123-134 + mcx_preprocess, src/mcx_utils.c:1688-1712): the 8 bytes of a voxel are
        # {lower label, upper label, px, py, pz, nx, ny, nz}; two planes of words come out, (2, x, y, z)
        b = v.astype(np.uint8).astype(np.uint32)
        hi = (b[0] << 24) | (b[1] << 16) | (b[2] << 8) | b[3]
        lo = (b[4] << 24) | (b[5] << 16) | (b[6] << 8) | b[7]
        return np.stack([hi, lo]), MEDIA_2LABEL_SPLIT
    if v.dtype in (np.int8, np.uint8) and ch == 4:
        b = v.astype(np.uint8).astype(np.uint32)
        return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24), MEDIA_ASGN_BYTE
    if v.dtype in (np.int16, np.uint16) and ch == 2:
        h = v.astype(np.uint16).astype(np.uint32)
        return h[0] | (h[1] << 16), MEDIA_AS_SHORT
    if v.dtype in (np.float32, np.float64):
        f = v.astype(np.float32)
        if ch == 1:
            mua = f[0] * u
            w = mua.view(np.uint32).copy()
            w[w == 0] = np.float32(1.19209290e-07).view(np.uint32)      # avoid being taken for a 0-label voxel
            w[np.isnan(f[0])] = 0
            return w, MEDIA_MUA_FLOAT
        if ch == 2:
            mua, mus = f[0] * u, f[1] * u
            w = _float_to_half_bits(mua) | (_float_to_half_bits(mus) << 16)
            w[w == 0] = 0x00010000
            w[np.isnan(mua) | np.isnan(mus)] = 0
            return w.astype(np.uint32), MEDIA_AS_F2H
        if ch == 3:
            val, slot, lab = f[0].copy(), f[1], f[2]
            if (slot < 0).any() or (slot >= 4).any() or (lab < 0).any():
                raise ConfigError(-4, "the 2nd volume must have an integer value between 0 and 3")
            val[(slot >= 1) & (slot <= 2)] *= u
            i = val.view(np.uint32).astype(np.int64)
            e = (i >> 23) & 0xFF
            tmp = np.where(e > 0x70, (e - 0x70) & 0x1F, 0)
            hval = ((((i >> 31) << 5) | tmp) << 10) | ((i >> 13) & 0x3FF)
            low = ((slot.astype(np.int64) & 0x3) << 14) | lab.astype(np.int64)
            return ((hval << 16) | low).astype(np.uint32), MEDIA_LABEL_HALF
    raise ConfigError(-4, "Invalid array for vol array.")


def prepare(cfg):
    """dict -> Prepared.  Mirrors parse_config + mcx_validatecfg + mcx_preprocess for the hot path."""
    if "vol" not in cfg or "prop" not in cfg:
        raise ConfigError(-4, "You must define 'vol' and 'prop' field.")
    p = Prepared()
    c = p.c
    c.abi_version = abi.ABI_VERSION
    p.session = str(cfg.get("session", ""))

    vol = np.asarray(cfg["vol"])
    c.mediaformat = int(cfg.get("mediaformat", 0))          # explicit format code with already packed uint32 words
    if vol.ndim == 4:
        vol, c.mediaformat = pack_continuous_volume(vol, float(cfg.get("unitinmm", 1.0)))
    second = None
    if c.mediaformat == MEDIA_2LABEL_SPLIT and vol.ndim == 4:
        vol, second = vol[0], vol[1]
    if vol.ndim != 3:
        raise ConfigError(-4, "the 'vol' field must be a 3D array of labels or a 4D array of optical properties")
    if vol.size == 0:
        raise ConfigError(-4, "the 'vol' field in the input structure can not be empty")
    nx, ny, nz = vol.shape
    flat = np.ascontiguousarray(vol.astype(np.uint32, copy=False).ravel(order="F")).copy()
    if second is not None:      # the second word of every voxel follows the dimxyz first words (src/mcx_utils.c:1706-1707)
        flat = np.concatenate([flat, np.ascontiguousarray(second.astype(np.uint32, copy=False).ravel(order="F"))])
    c.dimx, c.dimy, c.dimz = nx, ny, nz

    prop = np.array(cfg["prop"], dtype=np.float32).reshape(-1, 4).copy()
    if prop.shape[0] == 0:
        raise ConfigError(-4, "you must define the 'prop' field in the input structure")
    c.medianum = prop.shape[0]

    # --- defaults of mcx_initcfg -------------------------------------------------------------
    c.tstart = float(cfg.get("tstart", 0.0))
    c.tend = float(cfg.get("tend", 5e-9))
    c.tstep = float(cfg.get("tstep", 5e-9))
    c.nphoton = int(cfg.get("nphoton", 0))
    c.replaydet = int(cfg.get("replaydet", 0))
    seedval = cfg.get("seed", 0x623F9A9E)
    replayseed = None
    if isinstance(seedval, np.ndarray) and seedval.ndim == 2:
        # an array of saved RNG states = photon replay (src/pmcxcl.cpp:1019-1048): 16 bytes per photon
        sv = np.asarray(seedval)
        if sv.dtype == np.uint64 and sv.shape[1] == 2:
            replayseed = np.ascontiguousarray(sv).copy()
        else:
            sv = sv.astype(np.uint8, copy=False)
            if sv.shape[0] != 16:
                raise ConfigError(-6, "the row number of cfg.seed does not match RNG seed byte-length")
            replayseed = np.ascontiguousarray(sv.T).copy().view(np.uint64).reshape(-1, 2)
        c.seed = SEED_FROM_FILE
        cfg = dict(cfg, nphoton=replayseed.shape[0])
    else:
        c.seed = int(seedval)
    c.seed_skip = int(cfg.get("seed_skip", 0))
    c.isreflect = int(cfg.get("isreflect", 1))
    c.isnormalized = int(cfg.get("isnormalized", 1))
    c.issavedet = int(cfg.get("issavedet", 1))
    c.issave2pt = int(cfg.get("issave2pt", 1))
    c.unitinmm = float(cfg.get("unitinmm", 1.0))
    c.minenergy = float(cfg.get("minenergy", 0.0))
    c.savedetflag = parse_savedetflag(cfg.get("savedetflag", 0x5))
    c.gscatter = int(cfg.get("gscatter", 1000000000))
    c.isspecular = int(cfg.get("isspecular", 0))
    c.maxvoidstep = int(cfg.get("maxvoidstep", 1000))
    c.voidtime = int(cfg.get("voidtime", 1))
    c.maxdetphoton = int(cfg.get("maxdetphoton", 1000000))
    c.srcnum = int(cfg.get("srcnum", 1))
    c.srcid = int(cfg.get("srcid", 0))
    c.issaveseed = int(cfg.get("issaveseed", 0))
    c.issaveref = int(cfg.get("issaveref", 0))
    c.nthread = int(cfg.get("nthread", 0)) if not cfg.get("autopilot", 1) or "nthread" in cfg else 0
    c.nblocksize = int(cfg.get("nblocksize", 0)) if "nblocksize" in cfg else 0
    c.sched = int(cfg.get("sched", 0))
    acc = cfg.get("accum", 0)
    c.accum = {"f64": 0, "f32": 1}.get(acc, acc) if isinstance(acc, str) else int(acc)
    dbg = cfg.get("debuglevel", 0)
    if isinstance(dbg, str):
        dbg = sum(1 << "RMPT".find(ch) for ch in dbg.upper() if ch in "RMPT")
    c.debuglevel = int(dbg) | (abi.DEBUG_STATS if cfg.get("stats") else 0)
    ot = cfg.get("outputtype", "flux")
    if isinstance(ot, str):
        if ot.lower() not in OUTPUTTYPES:
            raise ConfigError(-6, "output type %r is outside the photon-transport hot path of this build" % ot)
        ot = OUTPUTTYPES[ot.lower()]
    c.outputtype = int(ot)
    if c.outputtype in REPLAY_OUTPUTS and replayseed is None:
        raise ConfigError(-6, "output type %r needs photon replay: pass the saved seeds as cfg['seed'] and cfg['detphotons']" % (ot,))
    st = cfg.get("srctype", "pencil")
    if isinstance(st, str):
        if st not in SRCTYPES:
            raise ConfigError(-6, "the specified source type is not supported")
        st = SRCTYPES.index(st)
    c.srctype = int(st)
    respin = int(cfg.get("respin", 1))
    if respin == 0:
        raise ConfigError(-1, "respin number can not be 0, check your -r/--repeat input or cfg.respin value")       # src/mcx_utils.c:1628-1630
    if respin < 0:
        raise ConfigError(-1, "negative respin is not supported by this build")
    if replayseed is not None and respin > 1:
        respin = 1                           # "respin is disabled in the replay mode" (src/mcx_utils.c:1633-1636)
    c.respin = respin
    c.maxjumpdebug = int(cfg.get("maxjumpdebug", 10000000))                  # src/mcx_utils.c:288

    # --- sources: N x 4 arrays define extra sources (src/pmcxcl.cpp:477-660) -------------------
    def rows(key, default, wdef):
        a = np.atleast_2d(np.asarray(cfg.get(key, default), dtype=np.float64))
        return np.stack([_f4(r, wdef) for r in a])

    srcpos = rows("srcpos", [0, 0, 0, 1], 1.0)
    nsrc = srcpos.shape[0]
    srcdir = rows("srcdir", [0, 0, 1, 0], 0.0)
    srcp1 = rows("srcparam1", [0, 0, 0, 0], 0.0)
    srcp2 = rows("srcparam2", [0, 0, 0, 0], 0.0)

    def bcast(a):
        return a if a.shape[0] == nsrc else np.repeat(a[:1], nsrc, axis=0)

    srcdir, srcp1, srcp2 = bcast(srcdir), bcast(srcp1), bcast(srcp2)
    issrcfrom0 = int(cfg.get("issrcfrom0", 0))
    if not issrcfrom0:                      # mcx_validatecfg: convert to 0-based grid coordinates
        srcpos[:, :3] -= 1.0

    if c.tstart > c.tend or c.tstep == 0.0:
        raise ConfigError(-6, "incorrect time gate settings")
    if c.tend <= c.tstart:
        raise ConfigError(-6, "field 'tend' must be greater than field 'tstart'")

    # mcx_preprocess: normalise direction (double precision like the C code)
    # (only the main source, :1524-1533; extra sources are taken as given)
    d = srcdir[0, :3].astype(np.float64)
    n = math.sqrt(float((d * d).sum()))
    if n < 1e-10:
        raise ConfigError(-4, "source initial direction vector can not have a length of 0")
    srcdir[0, :3] = (d * (1.0 / n)).astype(np.float32)

    if c.debuglevel & abi.DEBUG_MOVE_ONLY:  # src/mcx_utils.c:1552-1555
        c.issave2pt = 0
        c.issavedet = 0
    if c.debuglevel & 1:                    # MCX_DEBUG_RNG
        c.isnormalized = 0
        c.issavedet = 0

    bc = parse_bc(cfg.get("bc"))
    isbcdet = bool(bc[6:].any())
    for i in range(12):
        c.bc[i] = int(bc[i])

    if c.unitinmm != 1.0:
        u = np.float32(c.unitinmm)
        prop[1:, 0] *= u
        prop[1:, 1] *= u
    detpos = np.array(cfg.get("detpos", np.zeros((0, 4))), dtype=np.float32).reshape(-1, 4).copy()
    if not issrcfrom0 and len(detpos):
        detpos[:, :3] -= 1.0
    c.detnum = detpos.shape[0]
    if c.issavedet and c.detnum == 0 and not isbcdet:
        c.issavedet = 0
    if c.issavedet == 0:
        c.savedetflag = 0
    prop[prop[:, 1] == 0.0, 1] = HOST_EPS
    srcpos[srcpos[:, 3] == 0.0, 3] = 1.0
    if nsrc > 1:
        if c.srcnum > 1:
            raise ConfigError(-4, "simulating multiple sources currently can not be used with photon-sharing")
        if c.srcid > nsrc:
            raise ConfigError(-4, "srcid exceeds total defined source count")
    if c.srcnum > 1 and c.srctype != 5:
        raise ConfigError(-4, "photon sharing (srcnum>1) needs the 'pattern' source type")

    continuous = c.mediaformat > 4
    if continuous:
        # src/mcx_utils.c:1760-1766
        if c.mediaformat in (MEDIA_AS_F2H, MEDIA_MUA_FLOAT, MEDIA_AS_HALF) and c.medianum < 2:
            raise ConfigError(-4, "the 'prop' field must contain at least 2 rows for the requested media format")
        if c.mediaformat in (MEDIA_ASGN_BYTE, MEDIA_AS_SHORT) and c.medianum < 3:
            raise ConfigError(-4, "the 'prop' field must contain at least 3 rows for the requested media format")
    maxlabel = 0 if continuous else int((flat & MED_MASK).max())
    if c.medianum <= maxlabel:
        raise ConfigError(-4, "input media optical properties are less than the labels in the volume")
    if c.srctype in (5, 15) and cfg.get("srcpattern") is None:
        raise ConfigError(-4, "the 'srcpattern' field can not be empty when your 'srctype' is 'pattern'")

    # launch voxel / label of point-like sources, as raw uint bits in param2.z/.w (:1718-1749)
    if c.srctype <= 2 or c.srctype in (7, 11):
        for i in range(nsrc):
            x, y, z = srcpos[i, :3]
            if x < 0 or y < 0 or z < 0 or x >= nx or y >= ny or z >= nz:
                idx, lab = 0, 0
            else:
                idx = int(math.floor(z)) * (ny * nx) + int(math.floor(y)) * nx + int(math.floor(x))
                lab = int(flat[idx] & MED_MASK)
            srcp2[i, 2:4] = np.array([idx, lab], dtype=np.uint32).view(np.float32)

    if c.issavedet:
        p.det_voxels = maskdet(flat[:nx * ny * nz], (nx, ny, nz), detpos, svmc=(c.mediaformat == MEDIA_2LABEL_SPLIT))
    if c.issavedet and c.savedetflag == 0:
        c.savedetflag = 0x5
    # polarised light: the Mueller-matrix tables mcx_prep_polarized (src/mcx_utils.c:1483-1519) would compute from `polprop`
    # are taken as given here (cfg['smatrix']: polmedianum x 181 x 4), with prop[] already holding the matching mus / g
    c.omega = float(cfg.get("omega", 0.0))
    if cfg.get("smatrix") is not None:
        sm = np.ascontiguousarray(np.asarray(cfg["smatrix"], dtype=np.float32).reshape(-1, abi.NANGLES, 4)).copy()
        if sm.shape[0] + 1 != c.medianum:
            raise ConfigError(-6, "number of polprop and prop is not consistent")            # src/mcx_utils.c:1540-1542
        p.keep["smatrix"] = sm
        c.polmedianum = sm.shape[0]
        c.smatrix = sm.ctypes.data_as(C.POINTER(abi.F4))
        iquv = _f4(cfg.get("srciquv", [1.0, 0.0, 0.0, 0.0]))      # default of mcx_initcfg (src/mcx_utils.c:338-339)
        c.srciquv = abi.F4(*[float(t) for t in iquv])
        if c.issavedet:
            c.savedetflag |= 0x04 | 0x20 | 0x40 | 0x80       # P, V, W, I are forced in polarised runs (src/mcx_utils.c:1777-1781)
    else:
        c.savedetflag &= ~0x80              # no Stokes vector without polarised media (src/mcx_utils.c:1777-1781)
    if c.issaveref > 1:
        raise ConfigError(-4, "issaveref > 1 is outside the hot path of this build")

    # --- publish buffers ---------------------------------------------------------------------
    src = np.zeros((nsrc, 16), dtype=np.float32)
    src[:, 0:4], src[:, 4:8], src[:, 8:12], src[:, 12:16] = srcpos, srcdir, srcp1, srcp2
    C.memmove(C.byref(c.src), src[0].ctypes.data, 64)
    c.extrasrclen = nsrc - 1
    extra = np.ascontiguousarray(src[1:])
    p.keep.update(vol=flat, prop=prop, detpos=detpos, srcall=src, extra=extra)
    c.vol = flat.ctypes.data_as(C.POINTER(C.c_uint32))
    c.prop = prop.ctypes.data_as(C.POINTER(abi.F4))
    c.detpos = detpos.ctypes.data_as(C.POINTER(abi.F4)) if c.detnum else None
    c.srcdata = extra.ctypes.data_as(C.POINTER(abi.Source)) if nsrc > 1 else None
    for key, cnt, ptr in (("invcdf", "nphase", "invcdf"), ("angleinvcdf", "nangle", "angleinvcdf")):
        if cfg.get(key) is not None:
            tabv = np.ascontiguousarray(np.asarray(cfg[key], dtype=np.float32).ravel()).copy()
            p.keep[key] = tabv
            setattr(c, cnt, tabv.size)
            setattr(c, ptr, tabv.ctypes.data_as(C.POINTER(C.c_float)))
    if cfg.get("srcpattern") is not None:
        pat = np.asarray(cfg["srcpattern"], dtype=np.float32)
        pat = np.ascontiguousarray(pat.ravel(order="F")).copy()
        p.keep["srcpattern"] = pat
        c.srcpattern = pat.ctypes.data_as(C.POINTER(C.c_float))
        c.srcpattern_len = pat.size
        if c.srctype == 5 and pat.size < int(srcp1[0, 3]) * int(srcp2[0, 3]) * max(1, c.srcnum):
            raise ConfigError(-4, "srcpattern is smaller than srcparam1.w x srcparam2.w")
    if replayseed is not None:
        _replayinit(p, cfg, replayseed, prop)
    return p


def _replayinit(p, cfg, seeds, prop):
    """mcx_replayinit (src/mcx_utils.c:1355-1421): keep the records of the replayed detector (and source), derive each
    photon's detected weight and time of flight from its partial paths.  `prop` is already scaled by unitinmm."""
    c = p.c
    detps = cfg.get("detphotons")
    if detps is None:
        raise ConfigError(-6, "you give cfg.seed for replay, but did not specify cfg.detphotons.")
    detps = np.asarray(detps.get("data") if isinstance(detps, dict) else detps, dtype=np.float32)
    if detps.ndim != 2 or detps.shape[1] != seeds.shape[0]:
        raise ConfigError(-6, "the column numbers of detphotons and seed do not match")
    flag = parse_savedetflag(cfg.get("savedetflag", 0x5)) or 0x5
    hasdetid = flag & 1
    nmed = c.medianum - 1
    if (not hasdetid and c.detnum > 1) or not (flag >> 2 & 1):
        raise ConfigError(-6, "please rerun the baseline simulation and save detector ID (D) and partial-path (P) using cfg.savedetflag='dp'")
    offset = (flag >> 1 & 1) * nmed
    ids = detps[0].astype(np.int32) if hasdetid else np.ones(detps.shape[1], np.int32)
    keep = np.ones(detps.shape[1], bool)
    if c.replaydet > 0:
        keep &= (ids & 0xFFFF) == c.replaydet
    if c.srcid > 0:
        keep &= ((ids >> 16) & 0xFFFF) == c.srcid
    plen = detps[offset + hasdetid: offset + hasdetid + nmed].astype(np.float32)          # (nmed, N), voxel units
    weight = np.ones(detps.shape[1], np.float32)
    tof = np.zeros(detps.shape[1], np.float32)
    for j in range(nmed):
        weight *= np.exp(-prop[j + 1, 0] * plen[j]).astype(np.float32)
        tof += plen[j] * np.float32(c.unitinmm) * R_C0 * prop[j + 1, 3]
    keep &= ~((tof < np.float32(c.tstart)) | (tof > np.float32(c.tend)))
    rs = np.ascontiguousarray(seeds[keep])
    rw = np.ascontiguousarray(weight[keep])
    rt = np.ascontiguousarray(tof[keep])
    rd = np.ascontiguousarray(ids[keep].astype(np.int32))
    p.keep.update(replay_seed=rs, replay_weight=rw, replay_tof=rt, replay_detid=rd, replay_index=np.nonzero(keep)[0])
    c.nphoton = rs.shape[0]
    c.replay_seed = rs.ctypes.data_as(C.POINTER(C.c_uint64))
    c.replay_weight = rw.ctypes.data_as(C.POINTER(C.c_float))
    c.replay_tof = rt.ctypes.data_as(C.POINTER(C.c_float))
    c.replay_detid = rd.ctypes.data_as(C.POINTER(C.c_int32))
