"""Small simulation decks that exercise every branch of the photon-transport path (sources, boundary
conditions, gates, detectors).  The physics follows the reference's own test commands
(test/testmcx.sh:60-132) and example decks where one exists."""
import numpy as np

from mcxcl_b200 import abi, benchmarks, hostcfg


def cube(nphoton=1e5, **kw):
    cfg = benchmarks.cube60b(nphoton)
    cfg.update(kw)
    return cfg


def pattern8():
    pat = np.zeros((8, 8), dtype=np.float32)
    pat[1:7, 2] = 1
    pat[1:7, 5] = 1
    pat[3, 2:6] = 1
    pat[5, 3] = 0.5
    return pat


def pattern3d():
    rs = np.random.RandomState(3)
    return rs.uniform(0, 1, (6, 5, 4)).astype(np.float32)


SOURCES = {
    # name: overrides on top of cube60b
    "pencil": dict(),
    "isotropic": dict(srctype="isotropic", srcpos=[29.0, 29.0, 29.0]),                     # testmcx.sh:88-90
    "cone": dict(srctype="cone", srcpos=[29.0, 29.0, 0.0], srcparam1=[0.6, 0, 0, 0]),
    "cone_uniformangle": dict(srctype="cone", srcpos=[29.0, 29.0, 0.0], srcparam1=[0.6, 1, 0, 0]),
    "gaussian": dict(srctype="gaussian", srcpos=[29.0, 29.0, 0.0], srcparam1=[8.0, 0, 0, 0]),
    "gaussian_focused": dict(srctype="gaussian", srcpos=[29.0, 29.0, 0.0, 1.0], srcdir=[0, 0, 1, 20.0], srcparam1=[4.0, 0.005, 0, 0]),
    "planar": dict(srctype="planar", srcpos=[10.0, 10.0, -10.0], srcparam1=[40.0, 0, 0, 0], srcparam2=[0, 40.0, 0, 0]),
    "planar_inside": dict(srctype="planar", srcpos=[10.0, 10.0, 0.0], srcparam1=[40.0, 0, 0, 0], srcparam2=[0, 40.0, 0, 0]),
    "pattern": dict(srctype="pattern", srcpos=[10.0, 10.0, 0.0], srcparam1=[40.0, 0, 0, 8], srcparam2=[0, 40.0, 0, 8], srcpattern=pattern8()),
    "fourier": dict(srctype="fourier", srcpos=[10.0, 10.0, 0.0], srcparam1=[40.0, 0, 0, 2], srcparam2=[0, 40.0, 0, 1.25]),
    "arcsine": dict(srctype="arcsine", srcpos=[29.0, 29.0, 29.0]),
    "disk": dict(srctype="disk", srcpos=[29.0, 29.0, 0.0], srcparam1=[15.0, 0, 0, 0]),
    "disk_tilted": dict(srctype="disk", srcpos=[29.0, 29.0, 5.0], srcdir=[0.3, 0.2, 0.9327379], srcparam1=[6.0, 2.0, 0, 0]),
    "disk_sector_focus": dict(srctype="disk", srcpos=[29.0, 29.0, 0.0], srcdir=[0, 0, 1, 25.0], srcparam1=[12.0, 0, 0.3, 2.0]),
    "fourierx": dict(srctype="fourierx", srcpos=[10.0, 10.0, 0.0], srcparam1=[40.0, 0, 0, 40.0], srcparam2=[2.0, 1.0, 0.25, 0.1]),
    "fourierx2d": dict(srctype="fourierx2d", srcpos=[10.0, 10.0, 0.0], srcparam1=[40.0, 0, 0, 40.0], srcparam2=[2.0, 1.0, 0.25, 0.1]),
    "zgaussian": dict(srctype="zgaussian", srcpos=[29.0, 29.0, 0.0], srcparam1=[0.4, 0, 0, 0]),
    "line": dict(srctype="line", srcpos=[15.0, 29.0, 10.0], srcdir=[0, 0, 1], srcparam1=[30.0, 0, 0, 0]),
    "line_fan": dict(srctype="line", srcpos=[15.0, 29.0, 0.0], srcdir=[0, 0, 1], srcparam1=[30.0, 0, 0, 0], srcparam2=[0.5, 0, 0, 0]),
    "slit": dict(srctype="slit", srcpos=[15.0, 29.0, 0.0], srcdir=[0, 0, 1], srcparam1=[30.0, 0, 0, 0]),
    "slit_diverging": dict(srctype="slit", srcpos=[15.0, 29.0, 0.0], srcdir=[0, 0, 1], srcparam1=[30.0, 0, 0, 0], srcparam2=[0.2, 0.1, 0, 0]),
    "pencilarray": dict(srctype="pencilarray", srcpos=[10.0, 10.0, 0.0], srcparam1=[40.0, 0, 0, 4], srcparam2=[0, 40.0, 0, 5]),   # testmcx.sh:100-102
    "pattern3d": dict(srctype="pattern3d", srcpos=[20.0, 20.0, 2.0], srcparam1=[6, 5, 4, 0], srcpattern=pattern3d()),
    "hyperboloid": dict(srctype="hyperboloid", srcpos=[29.0, 29.0, 0.0], srcparam1=[5.0, 10.0, 20.0, 0]),
    "ring": dict(srctype="ring", srcpos=[29.0, 29.0, 0.0], srcparam1=[15.0, 10.0, 0, 0]),
    "pencil_isotropic_launch": dict(srcpos=[29.0, 29.0, 29.0], srcdir=[0, 0, 1, float("nan")]),
    "disk_lambertian": dict(srctype="disk", srcpos=[29.0, 29.0, 0.0], srcdir=[0, 0, 1, float("-inf")], srcparam1=[10.0, 0, 0, 0]),
    "planar_diverging": dict(srctype="planar", srcpos=[20.0, 20.0, 0.0], srcdir=[0, 0, 1, -15.0], srcparam1=[20.0, 0, 0, 0], srcparam2=[0, 20.0, 0, 0]),
    "planar_outside_oblique": dict(srctype="planar", srcpos=[5.0, 5.0, -20.0], srcdir=[0.2, 0.1, 0.9746794], srcparam1=[30.0, 0, 0, 0], srcparam2=[0, 30.0, 0, 0], isspecular=1),
}

MIRROR_PROP = [[0, 0, 1, 1], [0.005, 1.0, 0.01, 1.37], [0.002, 5.0, 0.9, 1.0], [0.005, 1.0, 0.01, 1.37]]

BOUNDARIES = {
    "absorb_all": dict(isreflect=0),
    "reflect_all": dict(isreflect=1),
    "aarraa": dict(isreflect=0, bc="aarraa"),                                      # testmcx.sh:72-74  -> 27.x%
    "cyclic": dict(isreflect=0, bc="cccccc", nphoton=2000, issavedet=0),           # testmcx.sh:76-78  -> 99.x%
    # mirror decks carry a 4th media row equal to row 1: after a reflection at an exterior face the reference
    # ORs the boundary code into the label of the current voxel (src/mcx_core.cl:2796, 2928: mediaidold =
    # mediaid | isdet, restored when the packet scatters inside the voxel), so label 1 reads row 1|3 = 3
    "mirror_sides": dict(isreflect=0, bc="mmaamm", prop=MIRROR_PROP),
    "mirror_top_reflect": dict(isreflect=1, bc="rrmrrr", prop=MIRROR_PROP),
    "detect_faces": dict(isreflect=1, bc="______111111", nphoton=10000),           # testmcx.sh:104-106 -> 97x-99x detected
    "detect_low_faces_absorb": dict(isreflect=0, bc="aaaaaa111000", nphoton=20000),
}


def two_layer(nphoton=1e5, **kw):
    """60^3 with three labels (mismatched indices between layers) and an air pocket: interior Fresnel faces."""
    vol = np.ones((60, 60, 60), dtype=np.uint8)
    vol[:, :, 10:25] = 2
    vol[:, :, 25:] = 3
    vol[20:40, 20:40, 30:40] = 0
    cfg = benchmarks.cube60b(nphoton)
    cfg.update(vol=vol, prop=[[0, 0, 1, 1], [0.005, 1.0, 0.01, 1.37], [0.02, 5.0, 0.9, 1.5], [0.001, 0.5, 0.8, 1.33]])
    cfg.update(kw)
    return cfg


def many_labels(nphoton=1e5, **kw):
    """more than 127 labels: exercises the 16-bit media path"""
    rs = np.random.RandomState(9)
    vol = (1 + (np.arange(60)[:, None, None] // 2 + 30 * (np.arange(60)[None, :, None] // 12) + 0 * np.arange(60)[None, None, :])).astype(np.uint16)
    nlab = int(vol.max()) + 1
    prop = np.zeros((nlab, 4), dtype=np.float32)
    prop[0] = [0, 0, 1, 1]
    prop[1:, 0] = rs.uniform(0.002, 0.02, nlab - 1)
    prop[1:, 1] = rs.uniform(0.5, 3.0, nlab - 1)
    prop[1:, 2] = rs.uniform(0.0, 0.9, nlab - 1)
    prop[1:, 3] = rs.choice([1.33, 1.37, 1.4], nlab - 1)
    cfg = benchmarks.cube60b(nphoton)
    cfg.update(vol=vol, prop=prop)
    cfg.update(kw)
    return cfg


# ---- continuous media (Config.mediabyte 99-104): volumes whose words encode the optical properties, packed by hostcfg like pmcxcl ----
MEDIA_PROP2 = [[0, 0, 1, 1], [0.005, 1.0, 0.01, 1.37]]
MEDIA_PROP3 = [[0, 0, 1, 1], [0.0, 0.0, 0.01, 1.0], [0.02, 5.0, 0.9, 1.5]]          # rows 1 and 2 = the range of the scaled formats


def media_two_regions():
    """mua / mus maps: a 60^3 cube with a more absorbing, more scattering slab at z = 20..39"""
    mua = np.full((60, 60, 60), 0.005, np.float32)
    mus = np.full((60, 60, 60), 1.0, np.float32)
    mua[:, :, 20:40] = 0.015
    mus[:, :, 20:40] = 2.5
    return mua, mus


def media_volumes():
    mua, mus = media_two_regions()
    out = {}
    out["mua_float"] = (mua[None], MEDIA_PROP2, hostcfg.MEDIA_MUA_FLOAT)
    out["as_f2h"] = (np.stack([mua, mus]), MEDIA_PROP2, hostcfg.MEDIA_AS_F2H)
    lh = np.zeros((3, 60, 60, 60), np.float32)            # {value, slot, label}: mua of label 1 replaced inside the slab
    lh[0], lh[1], lh[2] = mua, 0, 1
    out["label_half"] = (lh, MEDIA_PROP2, hostcfg.MEDIA_LABEL_HALF)
    b = np.zeros((4, 60, 60, 60), np.uint8)                # bytes scale between PROP3 rows 1 and 2
    b[0] = np.round(mua / 0.02 * 255)
    b[1] = np.round(mus / 5.0 * 255)
    b[2] = 0
    b[3] = 94                                              # n = 1 + 94/127 * 0.5 = 1.37
    b[3, :, :, 20:40] = 127                                # ... and 1.5 inside the slab: interior Fresnel faces
    out["asgn_byte"] = (b, MEDIA_PROP3, hostcfg.MEDIA_ASGN_BYTE)
    sh = np.zeros((2, 60, 60, 60), np.uint16)
    sh[0] = np.round(mua / 0.02 * 65535)
    sh[1] = np.round(mus / 5.0 * 65535)
    out["as_short"] = (sh, [[0, 0, 1, 1], [0.0, 0.0, 0.01, 1.37], [0.02, 5.0, 0.01, 1.37]], hostcfg.MEDIA_AS_SHORT)
    return out


# ---- polarised light: Mueller-matrix tables on the 181-point polar grid ----
def rayleigh(nmed):
    """Mueller matrix of a Rayleigh scatterer on the 181-point polar grid mcx_prep_polarized uses (src/mcx_utils.c:1483-1519):
    rows {S11, S12, S33, S43}"""
    c = np.cos(np.pi * np.arange(abi.NANGLES) / (abi.NANGLES - 1))
    m = np.stack([0.75 * (1 + c * c), -0.75 * (1 - c * c), 1.5 * c, 0 * c], axis=1).astype(np.float32)
    return np.repeat(m[None], nmed, axis=0)


def isotropic_matrix(nmed):
    m = np.zeros((abi.NANGLES, 4), np.float32)
    m[:, 0] = 1.0
    m[:, 2] = 1.0
    return np.repeat(m[None], nmed, axis=0)


def pol_cfg(n, **kw):
    cfg = dict(benchmarks.get("cube60b", n), prop=[[0, 0, 1, 1], [0.005, 1.0, 0.0, 1.37]], savedetflag="dpxvwi", smatrix=rayleigh(1), srciquv=[1, 1, 0, 0])
    cfg.update(kw)
    return cfg
