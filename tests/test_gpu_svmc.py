"""Split-voxel Monte Carlo media (MED_TYPE 97 = MEDIA_2LABEL_SPLIT).

Reference: updateproperty_svmc / ray_plane_intersect / reflectray_svmc src/mcx_core.cl:1231-1344 and their call sites in
the photon loop (:2638-2648, 2666-2669, 2686-2699, 2716-2747, 2778-2782, 2816-2825, 2931-2949, 2962-2966, 3074-3111,
3218-3244); host packing src/pmcxcl.cpp:123-134, src/mcx_utils.c:1688-1712, detector voxels :4157-4168.

The reference's tests hold no pin for this mode; the checker is the reference's kernel source built with -DMED_TYPE=97
(oracle/build_ref.py), run on the same decks."""
import numpy as np
import pytest

from mcxcl_b200 import engine, hostcfg


def tilted_slab(n=(40, 40, 40), z0=14.3, sx=0.25, sy=-0.15, below=1, above=2):
    """pmcxcl's 8-byte-per-voxel SVMC volume (8, nx, ny, nz): {lower label, upper label, px, py, pz, nx, ny, nz}.  Two
    tissues separated by the plane z = z0 + sx (x - nx/2) + sy (y - ny/2); the voxels it cuts become split voxels.  The
    "upper" part of a split voxel is the side the normal points to; an empty (label 0) part has to be the LOWER one
    (src/mcx_utils.c:4157-4168), so the normal points away from the background."""
    nx, ny, nz = n
    v = np.zeros((8, nx, ny, nz), np.uint8)
    up = 1.0 if above != 0 else -1.0                      # normal towards +z unless the region above is the background
    lower, upper = (below, above) if up > 0 else (above, below)
    nrm = up * np.array([-sx, -sy, 1.0])
    nrm /= np.linalg.norm(nrm)
    nb = np.clip(np.round((nrm + 1) * 255 / 2), 0, 255).astype(np.uint8)
    for ix in range(nx):
        for iy in range(ny):
            zc = [z0 + sx * (ix + a - nx / 2) + sy * (iy + b - ny / 2) for a in (0, 1) for b in (0, 1)]
            zmin, zmax = min(zc), max(zc)
            zp = z0 + sx * (ix + 0.5 - nx / 2) + sy * (iy + 0.5 - ny / 2)
            for iz in range(nz):
                if iz + 1 <= zmin:
                    v[0, ix, iy, iz] = below
                elif iz >= zmax:
                    v[0, ix, iy, iz] = above
                else:
                    fz = min(max(zp - iz, 0.0), 1.0)
                    v[:, ix, iy, iz] = [lower, upper, 128, 128, int(round(fz * 255)), nb[0], nb[1], nb[2]]
    return v


def deck(nphoton, **kw):
    cfg = dict(vol=tilted_slab(), prop=[[0, 0, 1, 1], [0.02, 1.0, 0.8, 1.37], [0.005, 2.0, 0.9, 1.37]], nphoton=nphoton, srcpos=[20, 20, 0],
               srcdir=[0, 0, 1], issrcfrom0=1, tstart=0, tend=5e-9, tstep=5e-9, isreflect=1, seed=12345, issavedet=0, isnormalized=0)
    cfg.update(kw)
    return cfg


def test_host_packing_of_split_voxels():
    """mcx_preprocess (src/mcx_utils.c:1688-1712): first word {lower, upper, px, py} from the top byte down, second word
    {pz, nx, ny, nz}; all first words, then all second words"""
    v = np.zeros((8, 2, 3, 4), np.uint8)
    v[:, 1, 2, 3] = [5, 7, 10, 20, 30, 40, 50, 60]
    p = hostcfg.prepare(dict(deck(10), vol=v, prop=[[0, 0, 1, 1]] * 8, srcpos=[0, 0, 0]))
    assert p.c.mediaformat == 97 and p.dims == (2, 3, 4) and p.keep["vol"].size == 48
    i = 3 * 6 + 2 * 2 + 1
    assert p.keep["vol"][i] == (5 << 24 | 7 << 16 | 10 << 8 | 20) and p.keep["vol"][24 + i] == (30 << 24 | 40 << 16 | 50 << 8 | 60)


def test_oracle_split_voxels_are_close_to_the_label_volume(ref):
    """the reference source built with -DMED_TYPE=97: a plane through split voxels gives nearly what the staircase of the
    label volume gives (the two differ only inside the cut voxels)"""
    cfg = deck(40000)
    sv = ref.run(hostcfg.prepare(cfg), 256, hostthreads=0)
    v = cfg["vol"]
    lab = np.where(v[1] > 0, np.where(v[4] > 127, 1, 2), v[0]).astype(np.uint8)
    lb = ref.run(hostcfg.prepare(dict(cfg, vol=lab)), 256, hostthreads=0)
    assert abs(sv["absorbed"] - lb["absorbed"]) < 0.02
    a, b = sv["field"].reshape(40, 40, 40).sum(axis=(1, 2)), lb["field"].reshape(40, 40, 40).sum(axis=(1, 2))
    assert np.abs(a[:20] / b[:20] - 1).max() < 0.12


def profiles(field):
    f = field.astype(np.float64).reshape(40, 40, 40)          # [z][y][x]
    return f.sum(axis=(1, 2)), f.sum(axis=(0, 2)), f.sum(axis=(0, 1))


# The reference's kernel reads media[idx1dold] with idx1dold == OUTSIDE_VOLUME in rare walks through split voxels with an
# index mismatch (src/mcx_core.cl:3212-3213 after a reflection whose "previous voxel" lies outside the grid; found with an
# AddressSanitizer build of the oracle: 250000 photons on 4096 work-items crash, the sizes below do not).  The mismatched deck
# therefore runs the checker at a size that is known to stay inside its buffers; a walk is a function of (seed, work-items,
# photons) only, so this holds on every machine.
@pytest.mark.gpu
@pytest.mark.parametrize("name,nr,work,over", [
    ("matched", 250000, 4096, {}),
    ("mismatched", 120000, 1024, dict(prop=[[0, 0, 1, 1], [0.02, 1.0, 0.8, 1.37], [0.005, 2.0, 0.9, 1.55]])),
    ("no_reflection", 250000, 4096, dict(isreflect=0)),
    ("tilted_surface", 250000, 4096, dict(vol=tilted_slab(z0=30.4, sx=0.2, sy=0.1, below=1, above=0), prop=[[0, 0, 1, 1], [0.01, 1.0, 0.5, 1.37]])),
])
def test_gpu_split_voxels_match_the_reference_source(ref, name, nr, work, over):
    ng = 1000000
    p = hostcfg.prepare(deck(ng, **over))
    with engine.Simulation(p) as sim:
        assert sim.kernel_name.endswith("/ext")
    g = engine.run_prepared(p)
    o = ref.run(hostcfg.prepare(deck(nr, **over)), work, hostthreads=0)
    assert g["energytot"] == ng
    sigma = 1.6 * np.sqrt(o["absorbed"] * (1 - o["absorbed"]) / nr)
    assert abs(g["absorbed"] - o["absorbed"]) < 4 * sigma, (g["absorbed"], o["absorbed"])
    for a, b in zip(profiles(g["field"] / ng), profiles(o["field"] / nr)):
        big = b > 0.02 * b.max()
        assert np.abs(a[big] / b[big] - 1).max() < (0.05 if nr >= 250000 else 0.08)
    assert g["field"].astype(np.float64).sum() / ng == pytest.approx(o["field"].astype(np.float64).sum() / nr, rel=0.01)


@pytest.mark.gpu
def test_gpu_split_voxels_detect_on_the_tilted_surface(ref):
    """detectors on a surface made of split voxels (the empty part is the lower one): records carry the partial paths of
    the tissue the packet was in, by the current part of the voxel (src/mcx_core.cl:2778-2782, 1555-1561)"""
    over = dict(vol=tilted_slab(z0=30.4, sx=0.2, sy=0.1, below=1, above=0), prop=[[0, 0, 1, 1], [0.01, 1.0, 0.5, 1.37]], issavedet=1,
                detpos=[[20, 20, 30.4, 4], [10, 20, 28.4, 3]], savedetflag="dsp", maxdetphoton=300000)
    ng, nr = 2000000, 400000
    p = hostcfg.prepare(deck(ng, **over))
    assert sum(p.det_voxels) > 20
    g = engine.run_prepared(p)
    o = ref.run(hostcfg.prepare(deck(nr, **over)), 4096, hostthreads=0)
    assert o["detected"] > 300
    rg, ro = g["detected"] / ng, o["detected"] / nr
    assert abs(rg - ro) < 5 * np.sqrt(ro / nr) + 5 * np.sqrt(ro / ng)
    dg, do = g["detp"], o["detp"]
    for det in (1, 2):
        a, b = dg[dg[:, 0] == det], do[do[:, 0] == det]
        assert len(b) > 50
        # mean partial path and mean scattering count of the detected packets
        se = b[:, 2].std() / np.sqrt(len(b)) + a[:, 2].std() / np.sqrt(len(a))
        assert abs(a[:, 2].mean() - b[:, 2].mean()) < 5 * se
        na, nb = a[:, 1:2].copy().view(np.uint32).astype(np.float64), b[:, 1:2].copy().view(np.uint32).astype(np.float64)
        assert abs(na.mean() - nb.mean()) < 5 * (nb.std() / np.sqrt(len(nb)) + na.std() / np.sqrt(len(na)))
