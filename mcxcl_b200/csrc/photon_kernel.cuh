/*
 * photon_kernel.cuh -- the persistent photon-transport kernel for sm_100a.
 *
 * Work done per photon packet (the work of the reference's mcx_main_loop, src/mcx_core.cl:2307-3307, and
 * launchnewphoton, :1466-2275 -- see SURVEY.md section 8(a) rows R1-R16): source sampling, scattering-length
 * draw, Henyey-Greenstein deflection, voxel ray-marching with Beer-Lambert attenuation, time-gated
 * fluence deposit, boundary handling (Fresnel / mirror / cyclic / absorb), Russian roulette, and
 * detected-photon capture.
 *
 * B200 design:
 *   - one persistent grid, a multiple of the SM count; threads pull photon batches from a per-GPU
 *     counter (no static per-thread budget, so long-lived photons do not leave a tail);
 *   - the media volume is a read-only byte (or 16-bit) volume with the detector flag in the top bit,
 *     fetched through the non-coherent path; the optical-property table, extra sources and detector
 *     list live in shared memory;
 *   - fluence goes out as fire-and-forget RED.ADD into an L2-resident accumulator volume (fp64
 *     accumulators by default: they make the reference's MAX_ACCUM shadow-buffer spill unnecessary);
 *   - detected photons are compacted with a warp-aggregated atomic (ballot + popc);
 *   - source type, reflection, detector capture, media word size and accumulator type are compile-time
 *     specialisations, the same axes the reference JIT-specialises on (src/mcx_host.cpp:857-971).
 */
#pragma once
#include <cuda_fp16.h>
#include "photon_device.cuh"

namespace mcxb {

/* same numbering as MCX_SRC_* (src/mcx_const.h:75-92) */
enum SrcType {
    srcPencil = 0, srcIsotropic, srcCone, srcGaussian, srcPlanar, srcPattern, srcFourier, srcArcsine, srcDisk,
    srcFourierX, srcFourierX2D, srcZGaussian, srcLine, srcSlit, srcPencilArray, srcPattern3D, srcHyperboloid, srcRing,
    srcAny = -1      /* decided at run time from SimParam::srctype */
};

struct SimParam {
    /* domain */
    uint32_t nx, ny, nz;
    uint32_t dimxy, dimxyz;
    uint32_t fieldlen;            /* dimxyz * maxgate * number of output volumes */
    float    fnx, fny, fnz;
    /* time window */
    float    twin0, twin1, Rtstep;
    float    oneoverc0;           /* R_C0 * unitinmm                  (src/mcx_host.cpp:512) */
    float    minaccumtime;        /* unitinmm * R_C0 * unitinmm       (src/mcx_host.cpp:515) */
    uint32_t maxgate;
    /* physics switches */
    float    minenergy;
    uint32_t gscatter;
    uint32_t doreflect, save2pt, outputtype, is2d;
    uint32_t isspecular, issaveref, issaveseed;
    int32_t  voidtime;
    uint32_t maxvoidstep;
    uint8_t  bc[12];
    /* sources */
    int32_t  srctype, srcid;
    uint32_t extrasrclen, srcnum;
    /* tables in shared memory: [medianum] media, then 4*(1+extrasrclen) source rows, then detnum detectors */
    uint32_t medianum, detnum, tablen;
    const float4* tables;
    uint32_t mediaformat;         /* continuous media (32-bit media words): format code 99..104, else 0 */
    uint32_t plainlaunch;         /* one pencil source inside a labelled voxel, no focal length / angle table: the launch is a copy */
    const float*  srcpattern;
    /* launch-angle / phase-function inverse CDF tables (src/mcx_core.cl:2102-2118, 2475-2482) */
    uint32_t nphase, nangle, ftablen;   /* ftablen = nphase+nangle rounded up to an even count */
    const float* invcdf;
    const float* angleinvcdf;
    /* detected photons */
    uint32_t savedetflag, partialdata, reclen, maxdetphoton;
    float*    detphoton;
    uint32_t* detcount;
    unsigned long long* seedout;
    /* photon budget */
    unsigned long long  nphoton;
    unsigned long long* counter;  /* dynamic scheduling: next unclaimed photon */
    uint32_t chunk;               /* photons claimed per refill */
    uint32_t threadphoton;        /* static scheduling (src/mcx_host.cpp:1011-1012) */
    int32_t  oddphoton;
    int32_t  sched;               /* 0 dynamic, 1 static */
    /* volumes */
    const void* media;
    void*       field;
    const uint32_t* seeds;        /* 4 words per thread */
    double*     energy;           /* {escaped, launched} */
    unsigned long long* stats;    /* optional {segments, deposits, scatters} counters, or NULL */
    /* photon replay (src/mcx_core.cl:1590-1596, 2567-2592, 2845-2858): per-photon RNG state, detected weight, time of
     * flight and detector of the photons being replayed; replayseed == NULL in a forward run */
    const unsigned long long* replayseed;
    const float* replayweight;
    const float* replaytof;
    const int*   replaydetid;
    int32_t      replaydet;
    uint32_t     nrepvol;         /* volumes per source in replay: detnum when replaydet == -1, else 1 */
    /* trajectory capture (`-D M`, generic kernels): one atomic counter, MCX_DEBUG_REC_LEN = 6 floats per record */
    float*       trajdata;        /* NULL = off */
    uint32_t*    trajcount;
    uint32_t     maxjumpdebug;
    uint32_t     idbase;          /* photons launched by earlier batches (respin): photon ids continue across batches */
    /* replicated accumulators: CTA b adds into copy (b % acccopies); the copies are summed by finalize_kernel */
    uint32_t     widedep;         /* common kernels: bit 0 = more than one gate, bit 1 = one volume per source */
    uint32_t     acccopies;
    unsigned long long accstride; /* elements between two copies (= fieldlen) */
    /* extended physics (EXT kernels only) */
    const float4* smatrix;        /* polarised light: maxpolmedia x kNAngles rows {S11, S12, S33, S43} (src/mcx_core.cl:801-812), or NULL */
    uint32_t     maxpolmedia;     /* media with a Mueller matrix (Config.polmedianum), 0 = unpolarised */
    float        s0i, s0q, s0u, s0v;   /* Stokes vector of a fresh packet (Config.srciquv, :1633-1638) */
    float        omega;           /* RF modulation angular frequency (Config.omega) */
    uint32_t     rfforward;       /* omega > 0 in a forward run: complex packet weights (:2750-2763, 2833-2841) */
    unsigned long long rfplane;   /* RF outputs: elements between the real and the imaginary volume (0 = no second plane) */
    uint32_t     adjfirstdet;     /* adjoint runs / srcid == -2: sources with an id above this are detectors launched as disks (:2154-2183); else 0xFFFFFFFF */
};

constexpr int kNAngles = 181;      /* NANGLES, src/mcx_const.h:67 */


/* per-photon state that survives across segments */
struct Photon {
    float px, py, pz, w;          /* position (voxel units) and packet weight */
    float vx, vy, vz;             /* direction */
    int   nscat;                  /* scattering events so far (the reference keeps it in v.w, with an EPS sentinel until the first draw, :2254) */
    float slen;                   /* remaining unitless scattering length (f.x) */
    float tof;                    /* time of flight in seconds (f.y) */
    int   ix, iy, iz, face;       /* current voxel and the face crossed last (flipdir) */
    uint32_t idx1d;
    uint32_t label;               /* media label of the current voxel */
    uint32_t detflag;             /* detector bit of the current voxel, or the boundary code after leaving the grid */
    float w0;                     /* weight at the last deposit */
    float pathlen;                /* path length inside the current voxel */
    float n1;                     /* refractive index of the medium the packet is coming from */
};

template <typename MediaT> struct MediaTraits;
template <> struct MediaTraits<uint8_t> {
    static constexpr uint32_t det = 0x80u, lab = 0x7Fu;
};
template <> struct MediaTraits<uint16_t> {
    static constexpr uint32_t det = 0x8000u, lab = 0x7FFFu;
};

template <> struct MediaTraits<uint32_t> {        /* continuous media: the word IS the medium (src/mcx_core.cl:509-514) */
    static constexpr uint32_t det = 0x80000000u, lab = 0x7FFFFFFFu;
};

/* optical properties {mua, mus, g, n} of a voxel.  Label media: row `word` of the table.  Continuous media
 * (MediaT = uint32_t): decoded from the word like updateproperty (src/mcx_core.cl:1079-1193); the fields a format
 * does not carry come from row 1 (the reference leaves them at what launchnewphoton loaded from gproperty[1], :2210)
 * and n from row 0 / row 1 for an empty / non-empty word. */
template <typename MediaT>
__device__ __forceinline__ float4 medium(const SimParam& P, const float4* __restrict__ tab, uint32_t word) {
    if (sizeof(MediaT) < 4) {
        return tab[word];
    }

    const uint32_t fmt = P.mediaformat;
    float4 pr = tab[1];
    const float nfill = tab[word != 0u].w;

    if (fmt == 101u) {                                     /* MEDIA_MUA_FLOAT */
        pr.x = fabsf(__uint_as_float(word));
        pr.w = nfill;
    } else if (fmt == 100u || fmt == 102u) {               /* MEDIA_AS_F2H, MEDIA_AS_HALF */
        pr.x = fabsf(__half2float(__ushort_as_half((unsigned short)(word & 0xFFFFu))));
        pr.y = fabsf(__half2float(__ushort_as_half((unsigned short)(word >> 16))));
        pr.w = nfill;
    } else if (fmt == 99u) {                               /* MEDIA_LABEL_HALF */
        pr = tab[word & 0x3FFFu];
        const float v = fabsf(__half2float(__ushort_as_half((unsigned short)(word >> 16))));
        const uint32_t slot = (word & 0xC000u) >> 14;
        pr.x = (slot == 0u) ? v : pr.x;
        pr.y = (slot == 1u) ? v : pr.y;
        pr.z = (slot == 2u) ? v : pr.z;
        pr.w = (slot == 3u) ? v : pr.w;
    } else if (fmt == 103u) {                              /* MEDIA_ASGN_BYTE */
        const float4 lo = tab[1], hi = tab[2];
        pr.x = (float)(word & 0xFFu) * (1.f / 255.f) * (hi.x - lo.x) + lo.x;
        pr.y = (float)((word >> 8) & 0xFFu) * (1.f / 255.f) * (hi.y - lo.y) + lo.y;
        pr.z = (float)((word >> 16) & 0xFFu) * (1.f / 255.f) * (hi.z - lo.z) + lo.z;
        pr.w = (float)((word >> 24) & 0xFFu) * (1.f / 127.f) * (hi.w - lo.w) + lo.w;
    } else if (fmt == 104u) {                              /* MEDIA_AS_SHORT */
        const float4 lo = tab[1], hi = tab[2];
        pr.x = (float)(word & 0xFFFFu) * (1.f / 65535.f) * (hi.x - lo.x) + lo.x;
        pr.y = (float)(word >> 16) * (1.f / 65535.f) * (hi.y - lo.y) + lo.y;
        pr.w = nfill;
    }

    return pr;
}

template <typename MediaT>
__device__ __forceinline__ void fetch_voxel(const MediaT* __restrict__ media, uint32_t idx, uint32_t& label, uint32_t& det) {
    const uint32_t m = __ldg(media + idx);
    label = m & MediaTraits<MediaT>::lab;
    det = (m & MediaTraits<MediaT>::det) ? kDetMask : 0u;
}

__device__ __forceinline__ bool inside(const SimParam& P, float x, float y, float z) {
    return !(x < 0.f || y < 0.f || z < 0.f || x >= P.fnx || y >= P.fny || z >= P.fnz);
}

__device__ __forceinline__ uint32_t linear_index(const SimParam& P, int ix, int iy, int iz) {
    return (uint32_t)(iz * (int)P.dimxy + iy * (int)P.nx + ix);
}

__device__ __forceinline__ bool voxel_in_grid(const SimParam& P, int ix, int iy, int iz) {
    /* the reference compares (ushort)id against the float dimension (:2801) */
    return (uint32_t)(ix & 0xFFFF) < P.nx && (uint32_t)(iy & 0xFFFF) < P.ny && (uint32_t)(iz & 0xFFFF) < P.nz;
}

/* ---------------------------------------------------------------------------------------------------
 * march a packet launched outside the volume (or inside a zero voxel) up to the first non-zero voxel
 * (src/mcx_core.cl:1350-1455).  Returns the linear index of the entry voxel or -1.
 * ------------------------------------------------------------------------------------------------- */
/* inlined in the common and the extended kernels: as a real call it made every launch from outside the volume pay the ABI
 * (digimouse 353.1 -> 338.1 ms at 3e7 photons, cube60b 261.6 -> 258.1; profiles/r2_sweep_enter_inline.log).  The generic
 * kernels keep the call: inlined, their register allocation got worse (cube60b gscatter deck 42.0 -> 43.1 ms).
 * -DMCXB_ENTER_CALL restores the call everywhere (A/B). */
template <typename MediaT>
__device__ __forceinline__ int enter_volume_body(const SimParam& P, const float4* __restrict__ tab, Photon& ph);
template <typename MediaT>
__device__ __noinline__ int enter_volume_call(const SimParam& P, const float4* __restrict__ tab, Photon& ph) {
    return enter_volume_body<MediaT>(P, tab, ph);
}
template <typename MediaT, bool CALL>
__device__ __forceinline__ int enter_volume(const SimParam& P, const float4* __restrict__ tab, Photon& ph) {
#ifdef MCXB_ENTER_CALL
    return enter_volume_call<MediaT>(P, tab, ph);
#else
    if constexpr (CALL) return enter_volume_call<MediaT>(P, tab, ph);
    else return enter_volume_body<MediaT>(P, tab, ph);
#endif
}
template <typename MediaT>
__device__ __forceinline__ int enter_volume_body(const SimParam& P, const float4* __restrict__ tab, Photon& ph) {
    const MediaT* __restrict__ media = static_cast<const MediaT*>(P.media);
    int count = 1;
    ph.ix = (int)(short)floorf(ph.px);
    ph.iy = (int)(short)floorf(ph.py);
    ph.iz = (int)(short)floorf(ph.pz);
    ph.face = -1;

    while (true) {
        if (voxel_in_grid(P, ph.ix, ph.iy, ph.iz)) {
            int idx = (int)linear_index(P, ph.ix, ph.iy, ph.iz);
            uint32_t lab, det;
            fetch_voxel(media, (uint32_t)idx, lab, det);

            if (lab) {
                /* step back one unit and walk voxel by voxel to the entry face (:1364-1398) */
                ph.px -= ph.vx;
                ph.py -= ph.vy;
                ph.pz -= ph.vz;
                ph.ix = (int)(short)floorf(ph.px);
                ph.iy = (int)(short)floorf(ph.py);
                ph.iz = (int)(short)floorf(ph.pz);
                ph.tof -= P.minaccumtime;
                idx = (int)linear_index(P, ph.ix, ph.iy, ph.iz);
                count = 0;

                while (true) {
                    bool in = voxel_in_grid(P, ph.ix, ph.iy, ph.iz);

                    if (in) {
                        fetch_voxel(media, (uint32_t)idx, lab, det);

                        if (lab) {
                            break;
                        }
                    }

                    const float dist = face_distance(ph.px, ph.py, ph.pz, ph.vx, ph.vy, ph.vz, ph.ix, ph.iy, ph.iz, ph.face);
                    ph.tof += P.minaccumtime * dist;
                    ph.px = advance(ph.px, dist, ph.vx);
                    ph.py = advance(ph.py, dist, ph.vy);
                    ph.pz = advance(ph.pz, dist, ph.vz);

                    if (ph.face == 0) {
                        ph.ix += (ph.vx > 0.f ? 1 : -1);
                    } else if (ph.face == 1) {
                        ph.iy += (ph.vy > 0.f ? 1 : -1);
                    } else {
                        ph.iz += (ph.vz > 0.f ? 1 : -1);
                    }

                    idx = (int)linear_index(P, ph.ix, ph.iy, ph.iz);

                    if (count++ > 3) {
                        break;
                    }
                }

                ph.tof = P.voidtime ? ph.tof : 0.f;

                /* the reference reads media[idx] again here without a bounds check (:1420-1429); after a
                 * failed refinement idx may lie outside the grid, so clamp the read instead */
                uint32_t elab = 0;

                if ((uint32_t)idx < P.dimxyz) {
                    fetch_voxel(media, (uint32_t)idx, elab, det);
                }

                const float nin = medium<MediaT>(P, tab, elab).w, nout = tab[0].w;

                if (P.isspecular && nin != nout) {
                    ph.w *= 1.f - fresnel(ph.vx, ph.vy, ph.vz, nout, nin, ph.face);

                    if (ph.w > kEps) {
                        refract(ph.vx, ph.vy, ph.vz, nout, nin, ph.face);
                    }
                }

                return idx;
            }
        }

        if ((ph.px < 0.f && ph.vx <= 0.f) || (ph.px >= P.fnx && ph.vx >= 0.f) ||
                (ph.py < 0.f && ph.vy <= 0.f) || (ph.py >= P.fny && ph.vy >= 0.f) ||
                (ph.pz < 0.f && ph.vz <= 0.f) || (ph.pz >= P.fnz && ph.vz >= 0.f)) {
            return -1;
        }

        ph.px += ph.vx;
        ph.py += ph.vy;
        ph.pz += ph.vz;
        ph.ix = (int)(short)floorf(ph.px);
        ph.iy = (int)(short)floorf(ph.py);
        ph.iz = (int)(short)floorf(ph.pz);
        ph.tof += P.minaccumtime;

        if ((uint32_t)count++ > P.maxvoidstep) {
            return -1;
        }
    }
}

/* launch-time media lookup shared by the area sources (:1752-1758) */
template <typename MediaT>
__device__ __forceinline__ void locate(const SimParam& P, Photon& ph, uint32_t& rawlabel, uint32_t& rawdet) {
    ph.idx1d = (uint32_t)((int)floorf(ph.pz) * (int)P.dimxy + (int)floorf(ph.py) * (int)P.nx + (int)floorf(ph.px));

    if (inside(P, ph.px, ph.py, ph.pz)) {
        fetch_voxel(static_cast<const MediaT*>(P.media), ph.idx1d, rawlabel, rawdet);
    } else {
        rawlabel = 0;
        rawdet = 0;
    }
}

/* ---------------------------------------------------------------------------------------------------
 * sample one packet from the source description S[0..3] = {pos, dir, param1, param2}
 * (src/mcx_core.cl:1626-2153).  `focus` receives the point that the focal-length rule of :2140-2151
 * aims at; `aimed` is false for the sources that set their own direction (Lmove = 0 in the reference).
 * ------------------------------------------------------------------------------------------------- */
template <int SRC, typename MediaT>
__device__ __forceinline__ void sample_source(const SimParam& P, const float4* __restrict__ S, Rng& rng, Photon& ph,
        uint32_t& rawlabel, uint32_t& rawdet, float& fx, float& fy, float& fz, bool& aimed, uint32_t& patidx) {
    const int st = (SRC == srcAny) ? P.srctype : SRC;
    const float4 pos = S[0], dir = S[1], p1 = S[2], p2 = S[3];
    ph.px = pos.x;
    ph.py = pos.y;
    ph.pz = pos.z;
    ph.w = pos.w;
    ph.vx = dir.x;
    ph.vy = dir.y;
    ph.vz = dir.z;
    ph.idx1d = __float_as_uint(p2.z);
    rawlabel = __float_as_uint(p2.w) & 0x7FFFFFFFu;
    rawdet = __float_as_uint(p2.w) & kDetMask;
    fx = pos.x;
    fy = pos.y;
    fz = pos.z;
    aimed = true;

    if (st == srcPencil) {
        /* position, direction and launch voxel come straight from the source record */
    } else if (st == srcPlanar || st == srcPattern || st == srcPattern3D || st == srcFourier || st == srcPencilArray) {
        const float rx = rng_uniform(rng);
        const float ry = rng_uniform(rng);
        float rz = 0.f;

        if (st == srcPattern3D) {
            rz = rng_uniform(rng);
            ph.px += rx * p1.x;
            ph.py += ry * p1.y;
            ph.pz += rz * p1.z;
        }

        /* reference quirk, kept for parity: in the OpenCL build the `else` that separates the pattern3d offset
         * from the planar one is compiled only for __NVCC__ (:1675-1678), so a pattern3d packet receives the
         * planar offset rx*param1 + ry*param2 on top of its own ("reference quirks") */
        ph.px += rx * p1.x + ry * p2.x;
        ph.py += rx * p1.y + ry * p2.y;
        ph.pz += rx * p1.z + ry * p2.z;

        if (st == srcPattern) {
            const int cell = (int)(ry * kJustBelowOne * p2.w) * (int)p1.w + (int)(rx * kJustBelowOne * p1.w);

            if (P.srcnum > 1) {
                /* photon sharing (:1694-1705): one packet of unit weight carries every pattern; the deposit scales it
                 * by the srcnum pattern values of the launch cell */
                patidx = (uint32_t)cell;
                ph.w = 1.f;
            } else {
                ph.w = pos.w * __ldg(P.srcpattern + cell);
            }
        } else if (st == srcPattern3D) {
            ph.w = pos.w * __ldg(P.srcpattern + (int)(rz * kJustBelowOne * p1.z) * (int)p1.y * (int)p1.x
                                 + (int)(ry * kJustBelowOne * p1.y) * (int)p1.x + (int)(rx * kJustBelowOne * p1.x));
        } else if (st == srcFourier) {
            const float kx = floorf(p1.w), ky = floorf(p2.w);
            ph.w = pos.w * (__cosf((kx * rx + ky * ry + p1.w - kx) * kTwoPi) * (1.f - p2.w + ky) + 1.f) * 0.5f;
        } else if (st == srcPencilArray) {
            const float gx = floorf(rx * p1.w), gy = floorf(ry * p2.w);
            ph.px = pos.x + gx * p1.x / (p1.w - 1.f) + gy * p2.x / (p2.w - 1.f);
            ph.py = pos.y + gx * p1.y / (p1.w - 1.f) + gy * p2.y / (p2.w - 1.f);
            ph.pz = pos.z + gx * p1.z / (p1.w - 1.f) + gy * p2.z / (p2.w - 1.f);
        }

        locate<MediaT>(P, ph, rawlabel, rawdet);
        fx += (p1.x + p2.x) * 0.5f;
        fy += (p1.y + p2.y) * 0.5f;
        fz += (p1.z + p2.z) * 0.5f;
    } else if (st == srcFourierX || st == srcFourierX2D) {
        const float rx = rng_uniform(rng);
        const float ry = rng_uniform(rng);
        const float s = p1.w * fast_rsqrt(p1.x * p1.x + p1.y * p1.y + p1.z * p1.z);
        const float ux = s * (dir.y * p1.z - dir.z * p1.y);
        const float uy = s * (dir.z * p1.x - dir.x * p1.z);
        const float uz = s * (dir.x * p1.y - dir.y * p1.x);
        ph.px += rx * p1.x + ry * ux;
        ph.py += rx * p1.y + ry * uy;
        ph.pz += rx * p1.z + ry * uz;

        if (st == srcFourierX2D) {
            ph.w = pos.w * (__sinf((p2.x * rx + p2.z) * kTwoPi) * __sinf((p2.y * ry + p2.w) * kTwoPi) + 1.f) * 0.5f;
        } else {
            ph.w = pos.w * (__cosf((p2.x * rx + p2.y * ry + p2.z) * kTwoPi) * (1.f - p2.w) + 1.f) * 0.5f;
        }

        locate<MediaT>(P, ph, rawlabel, rawdet);
        fx += (p1.x + ux) * 0.5f;
        fy += (p1.y + uy) * 0.5f;
        fz += (p1.z + uz) * 0.5f;
    } else if (st == srcHyperboloid) {
        float sphi, cphi;
        __sincosf(kTwoPi * rng_uniform(rng), &sphi, &cphi);
        const float r = fast_sqrt(0.5f * rng_scatlen(rng)) * p1.x;
        const float k = -p1.y / p1.z;
        const float q = fast_rsqrt(r * r + p1.z * p1.z);
        float lx = r * (cphi - k * sphi), ly = r * (sphi + k * cphi), lz = 0.f;   /* local launch point */
        const float dx = -r * sphi * q, dy = r * cphi * q, dz = p1.z * q;         /* local direction    */

        if (ph.vz > -1.f + kEps && ph.vz < 1.f - kEps) {
            const float t = 1.f - ph.vz * ph.vz;
            const float st0 = fast_sqrt(t), rs = fast_rsqrt(t);
            const float cp = ph.vx * rs, sp = ph.vy * rs;
            const float gx = lx * cp * ph.vz - ly * sp, gy = lx * sp * ph.vz + ly * cp, gz = -lx * st0;
            const float nvx = dx * cp * ph.vz - dy * sp + dz * cp * st0;
            const float nvy = dx * sp * ph.vz + dy * cp + dz * sp * st0;
            const float nvz = -dx * st0 + dz * ph.vz;
            lx = gx;
            ly = gy;
            lz = gz;
            ph.vx = nvx;
            ph.vy = nvy;
            ph.vz = nvz;
        } else {
            const float s = (ph.vz > 0.f) ? dz : -dz;
            ph.vx = dx;
            ph.vy = dy;
            ph.vz = s;
        }

        ph.px = lx + pos.x;
        ph.py = ly + pos.y;
        ph.pz = lz + pos.z;
        aimed = false;
    } else if (st == srcDisk || st == srcGaussian || st == srcRing) {
        float phi;

        if (st != srcGaussian && (p1.z > 0.f || p1.w > 0.f)) {
            phi = fabsf(p1.z - p1.w) * rng_uniform(rng) + fminf(p1.z, p1.w);
        } else {
            phi = kTwoPi * rng_uniform(rng);
        }

        float sphi, cphi, r;
        __sincosf(phi, &sphi, &cphi);

        if (st != srcGaussian) {
            r = fast_sqrt(rng_uniform(rng) * fabsf(p1.x * p1.x - p1.y * p1.y) + p1.y * p1.y);
        } else if (fabsf(dir.w) < 1e-5f || fabsf(p1.y) < 1e-5f) {
            r = fast_sqrt(-0.5f * __logf(rng_uniform(rng))) * p1.x;
        } else {
            const float z0 = p1.x * p1.x * kOnePi / p1.y;
            r = fast_sqrt(-0.5f * __logf(rng_uniform(rng)) * (1.f + (dir.w * dir.w / (z0 * z0)))) * p1.x;
        }

        if (ph.vz > -1.f + kEps && ph.vz < 1.f - kEps) {
            const float t0 = 1.f - ph.vz * ph.vz;
            const float t1 = r * fast_rsqrt(t0);
            ph.px += t1 * (ph.vx * ph.vz * cphi - ph.vy * sphi);
            ph.py += t1 * (ph.vy * ph.vz * cphi + ph.vx * sphi);
            ph.pz -= t1 * t0 * cphi;
        } else {
            ph.px += r * cphi;
            ph.py += r * sphi;
        }

        locate<MediaT>(P, ph, rawlabel, rawdet);
    } else if (st == srcCone || st == srcIsotropic || st == srcArcsine) {
        float sphi, cphi, stheta, ctheta, ang;
        __sincosf(kTwoPi * rng_uniform(rng), &sphi, &cphi);

        if (st == srcCone) {
            do {
                ang = (p1.y > 0.f) ? kTwoPi * rng_uniform(rng) : acosf(2.f * rng_uniform(rng) - 1.f);
            } while (ang > p1.x);
        } else if (st == srcIsotropic) {
            ang = acosf(2.f * rng_uniform(rng) - 1.f);
        } else {
            ang = kOnePi * rng_uniform(rng);
        }

        __sincosf(ang, &stheta, &ctheta);
        rotate_direction(ph.vx, ph.vy, ph.vz, stheta, ctheta, sphi, cphi);
        aimed = false;
    } else if (st == srcZGaussian) {
        float sphi, cphi, stheta, ctheta;
        __sincosf(kTwoPi * rng_uniform(rng), &sphi, &cphi);
        const float a = fast_sqrt(-2.f * __logf(rng_uniform(rng)));
        const float ang = a * (1.f - 2.f * rng_uniform(rng)) * p1.x;
        __sincosf(ang, &stheta, &ctheta);
        rotate_direction(ph.vx, ph.vy, ph.vz, stheta, ctheta, sphi, cphi);
        aimed = false;
    } else if (st == srcLine || st == srcSlit) {
        float r = rng_uniform(rng);
        ph.px += r * p1.x;
        ph.py += r * p1.y;
        ph.pz += r * p1.z;

        if (st == srcLine) {
            float s, c;
            r = fast_rsqrt(p1.x * p1.x + p1.y * p1.y + p1.z * p1.z);

            if (p2.x > 0.f) {
                const float ax = p1.x * r, ay = p1.y * r, az = p1.z * r;
                const float d = ph.vx * ax + ph.vy * ay + ph.vz * az;
                ph.vx -= d * ax;
                ph.vy -= d * ay;
                ph.vz -= d * az;
                const float nrm = fast_rsqrt(ph.vx * ph.vx + ph.vy * ph.vy + ph.vz * ph.vz);
                ph.vx *= nrm;
                ph.vy *= nrm;
                ph.vz *= nrm;
                __sincosf(p2.x * (2.f * rng_uniform(rng) - 1.f), &s, &c);
                rotate_about_axis(ph.vx, ph.vy, ph.vz, ax, ay, az, s, c);
            } else {
                ph.vx = p1.x * r;
                ph.vy = p1.y * r;
                ph.vz = p1.z * r;
                __sincosf(kTwoPi * rng_uniform(rng), &s, &c);
                rotate_direction(ph.vx, ph.vy, ph.vz, 1.f, 0.f, s, c);
            }
        } else if (p2.x > 0.f || p2.y > 0.f) {
            float s, c;
            __sincosf(kTwoPi * rng_uniform(rng), &s, &c);
            r = fast_sqrt(2.f * rng_scatlen(rng));
            c *= p2.x * r;
            s *= p2.y * r;
            s *= fast_rsqrt(p1.x * p1.x + p1.y * p1.y + p1.z * p1.z);
            const float qx = p1.y * ph.vz - p1.z * ph.vy, qy = p1.z * ph.vx - p1.x * ph.vz, qz = p1.x * ph.vy - p1.y * ph.vx;
            c *= fast_rsqrt(qx * qx + qy * qy + qz * qz);
            ph.vx += c * qx + s * p1.x;
            ph.vy += c * qy + s * p1.y;
            ph.vz += c * qz + s * p1.z;
            r = fast_rsqrt(ph.vx * ph.vx + ph.vy * ph.vy + ph.vz * ph.vz);
            ph.vx *= r;
            ph.vy *= r;
            ph.vz *= r;
        }

        locate<MediaT>(P, ph, rawlabel, rawdet);
        fx = pos.x + p1.x * 0.5f;
        fy = pos.y + p1.y * 0.5f;
        fz = pos.z + p1.z * 0.5f;
        /* both the line and the slit source leave the focal-length rule enabled (:2082) */
    }
}

/* ---------------------------------------------------------------------------------------------------
 * split-voxel media (SVMC, MED_TYPE 97 = MEDIA_2LABEL_SPLIT, src/mcx_core.cl:1231-1344).  Two words per voxel:
 *   media[idx]          = lower label << 24 | upper label << 16 | px << 8 | py      (bit 31: detector flag)
 *   media[idx + dimxyz] = pz << 24 | nx << 16 | ny << 8 | nz
 * an oriented plane through the point (px, py, pz) / 255 inside the voxel with normal (nx, ny, nz) * 2/255 - 1 splits the
 * voxel into a "lower" and an "upper" (the side the normal points to) tissue; upper label 0 = an ordinary voxel.
 * The state keeps the plane oriented TOWARDS the other side of the packet, so only rays with v.n > 0 can hit it.
 * ------------------------------------------------------------------------------------------------- */
struct SplitVoxel {
    float nx, ny, nz, pd;      /* plane n.x = pd (nuvox.nv, nuvox.pd) */
    uint32_t sv;               /* bits 0-7 lower label, 8-15 upper label, 16 split, 17 packet is in the upper part (:577-587) */
};
__device__ __forceinline__ uint32_t sv_lower(uint32_t sv) {
    return sv & 0xFFu;
}
__device__ __forceinline__ uint32_t sv_upper(uint32_t sv) {
    return (sv >> 8) & 0xFFu;
}
__device__ __forceinline__ uint32_t sv_label(uint32_t sv) {
    return (sv & 0x20000u) ? sv_upper(sv) : sv_lower(sv);
}
__device__ __forceinline__ float dot3_rn(float ax, float ay, float az, float bx, float by, float bz) {
    return __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
}

/* updateproperty_svmc (:1231-1278): properties of the part of voxel idx1d the point p lies in, and the plane state */
__device__ __forceinline__ float4 svmc_update(const SimParam& P, const float4* __restrict__ tab, uint32_t word, uint32_t idx1d, float px, float py,
        float pz, int ix, int iy, int iz, SplitVoxel& nu) {
    if (idx1d == kOutsideMin || idx1d == kOutsideMax) {
        return tab[0];          /* the plane state is left as it was (:1235-1238) */
    }

    const uint32_t lo = __ldg(static_cast<const uint32_t*>(P.media) + idx1d + P.dimxyz);
    const uint32_t hi = word & 0x7FFFFFFFu;
    uint32_t sv = (hi >> 24) | (((hi >> 16) & 0xFFu) << 8);

    if ((hi >> 16) & 0xFFu) {
        const float rx = __fadd_rn(__fmul_rn((float)((hi >> 8) & 0xFFu), 1.f / 255.f), (float)ix);
        const float ry = __fadd_rn(__fmul_rn((float)(hi & 0xFFu), 1.f / 255.f), (float)iy);
        const float rz = __fadd_rn(__fmul_rn((float)(lo >> 24), 1.f / 255.f), (float)iz);
        float nx = __fsub_rn(__fmul_rn((float)((lo >> 16) & 0xFFu), 2.f / 255.f), 1.f);
        float ny = __fsub_rn(__fmul_rn((float)((lo >> 8) & 0xFFu), 2.f / 255.f), 1.f);
        float nz = __fsub_rn(__fmul_rn((float)(lo & 0xFFu), 2.f / 255.f), 1.f);
        const float r = rsqrtf(dot3_rn(nx, ny, nz, nx, ny, nz));
        nx = __fmul_rn(nx, r);
        ny = __fmul_rn(ny, r);
        nz = __fmul_rn(nz, r);
        float pd = dot3_rn(rx, ry, rz, nx, ny, nz);
        float4 pr;

        if (dot3_rn(px, py, pz, nx, ny, nz) > pd) {
            pr = tab[sv_upper(sv)];
            sv |= 0x20000u;
            nx = -nx;
            ny = -ny;
            nz = -nz;
            pd = -pd;
        } else {
            pr = tab[sv_lower(sv)];
        }

        nu.nx = nx;
        nu.ny = ny;
        nu.nz = nz;
        nu.pd = pd;
        nu.sv = sv | 0x10000u;
        return pr;
    }

    nu.sv = sv & 0xFFFFu;       /* an ordinary voxel; the plane coefficients keep their last values */
    return tab[hi >> 24];
}

/* one trajectory record (savedebugdata, src/mcx_core.cl:929-948): {photon id, x, y, z, weight, source id} */
static __device__ __noinline__ void save_traj(const SimParam& P, uint32_t id, float x, float y, float z, float w, int srcid) {
    const uint32_t pos = atomicAdd(P.trajcount, 1u);

    if (pos < P.maxjumpdebug) {
        float* rec = P.trajdata + 6 * (size_t)pos;
        rec[0] = __uint_as_float(id);
        rec[1] = x;
        rec[2] = y;
        rec[3] = z;
        rec[4] = w;
        rec[5] = (float)srcid;
    }
}

/* ---------------------------------------------------------------------------------------------------
 * detected-photon record (src/mcx_core.cl:838-926), compacted with one atomic per converged warp
 * ------------------------------------------------------------------------------------------------- */
__device__ __forceinline__ void save_detected(const SimParam& P, const float4* __restrict__ dets, const float* ppath,
        uint32_t pstride, const Photon& ph, uint32_t detarg, float w0init, int cursrc,
        const unsigned long long* photonseed, const float* stokes = nullptr) {
    int detid = 0;

    if (detarg == kOutsideMin) {
        detid = -1;
    } else {
        for (uint32_t i = 0; i < P.detnum; i++) {
            const float4 d = dets[i];
            const float dx = d.x - ph.px, dy = d.y - ph.py, dz = d.z - ph.pz;

            if (dx * dx + dy * dy + dz * dz < d.w * d.w) {
                detid = (int)i + 1;
                break;
            }
        }
    }

    const uint32_t active = __activemask();
    const uint32_t hits = __ballot_sync(active, detid != 0);

    if (detid == 0) {
        return;
    }

    const uint32_t leader = __ffs(hits) - 1;
    uint32_t base = 0;

    if (lane_id() == leader) {
        base = atomicAdd(P.detcount, (uint32_t)__popc(hits));
    }

    base = __shfl_sync(hits, base, leader) + __popc(hits & lanemask_lt());

    if (base >= P.maxdetphoton) {
        return;     /* overflow: counted but not stored, the host warns like src/mcx_host.cpp:1207-1210 */
    }

    if (P.issaveseed) {
        P.seedout[2 * (size_t)base] = photonseed[0];
        P.seedout[2 * (size_t)base + 1] = photonseed[pstride];
    }

    float* rec = P.detphoton + (size_t)base * P.reclen;
    const uint32_t flag = P.savedetflag;

    if (flag & 0x01u) {
        if (P.extrasrclen && P.srcid <= 0) {
            detid |= cursrc << 16;
        }

        *rec++ = (float)detid;
    }

    for (uint32_t i = 0; i < P.partialdata; i++) {
        *rec++ = ppath[i * pstride];
    }

    if (flag & 0x10u) {
        *rec++ = ph.px;
        *rec++ = ph.py;
        *rec++ = ph.pz;
    }

    if (flag & 0x20u) {
        *rec++ = ph.vx;
        *rec++ = ph.vy;
        *rec++ = ph.vz;
    }

    if (flag & 0x40u) {
        *rec++ = w0init;
    }

    if ((flag & 0x80u) && stokes) {       /* SAVE_IQUV (:912-917) */
        *rec++ = stokes[0];
        *rec++ = stokes[1];
        *rec++ = stokes[2];
        *rec++ = stokes[3];
    }
}

/* Stokes vector after a scattering event by polar angle theta and azimuth phi (updatestokes, src/mcx_core.cl:801-835):
 * rotate into the scattering plane, apply the Mueller matrix row of this medium and angle, rotate into the new
 * meridian plane, normalise to I = 1.  (ux,uy,uz) / (nx,ny,nz) = direction before / after the event. */
__device__ __forceinline__ void update_stokes(float& sI, float& sQ, float& sU, float& sV, float theta, float costheta, float phi, float s2p, float c2p,
        float uz, float nz, const float4* __restrict__ sm, uint32_t label) {
    /* costheta and sin / cos(2 phi) come from the sampling step that drew the angles (the reference recomputes them) */
    const float qi = sQ * c2p + sU * s2p, ui = -sQ * s2p + sU * c2p;
    const float4 m = __ldg(sm + (size_t)kNAngles * (label - 1u) + (uint32_t)(theta * (float)kNAngles * (0.318309886183791f - kEps)));
    const float i1 = m.x * sI + m.y * qi, q1 = m.y * sI + m.x * qi, u1 = m.z * ui + m.w * sV, v1 = -m.w * ui + m.z * sV;
    const float temp = (nz > -1.f && nz < 1.f) ? fast_rsqrt((1.f - costheta * costheta) * (1.f - nz * nz)) : 0.f;
    float cosi = (temp == 0.f) ? 0.f : (((phi > kOnePi && phi < kTwoPi) ? 1.f : -1.f) * (nz * costheta - uz) * temp);
    cosi = fmaxf(-1.f, fminf(cosi, 1.f));
    const float sini = fast_sqrt(1.f - cosi * cosi);
    const float cos22 = 2.f * cosi * cosi - 1.f, sin22 = 2.f * sini * cosi;
    const float inv = mufu_rcp(i1);
    sI = 1.f;
    sQ = (q1 * cos22 - u1 * sin22) * inv;
    sU = (q1 * sin22 + u1 * cos22) * inv;
    sV = v1 * inv;
}

/* ---------------------------------------------------------------------------------------------------
 * the kernel
 *
 * Compile-time axes (the reference JIT-specialises on the same ones, src/mcx_host.cpp:857-971):
 *   SRC      source type, or srcAny = decided at run time
 *   REFLECT  index-mismatch handling compiled in (MCX_DO_REFLECTION)
 *   SAVEDET  detected-photon capture (MCX_SAVE_DETECTORS): 0 = compiled out, 1 = the default record "DP" (detector id +
 *            partial paths, src/mcx_utils.c:266) folded at compile time, 2 = record flags read at run time
 *   MediaT   uint8_t (<=127 labels) or uint16_t media words
 *   AccT     double or float fluence accumulators
 *   STATS    count segments / deposits / scattering events (instrumented build, used to measure SURVEY 8(d))
 *   QDEPTH   depth of the per-thread scattering queue (see above), 0 = scatter in place
 *   GEN      false = the common configuration, with everything below decided at compile time:
 *              3-D domain, Henyey-Greenstein phase function, no gscatter switch, flux or fluence
 *              output with save2pt on, no diffuse-reflectance output, and -- unless BCODES is set -- all six boundary
 *              codes "unknown" (i.e. governed by isreflect alone) and no detect-on-face flags;
 *   BCODES   common kernels that honour per-face boundary codes and detect-on-face flags (resolved in the tail block) and
 *            also serve the energy and path-length output types (one uniform select in the deposit);
 *            true  = every option read from SimParam at run time.
 * ------------------------------------------------------------------------------------------------- */
#ifndef MCXB_BLOCK
    #define MCXB_BLOCK 256
#endif
constexpr int kBlock = MCXB_BLOCK;
#ifndef MCXB_MINBLOCKS
    #define MCXB_MINBLOCKS 4      /* 64 registers/thread, 32 resident warps per SM: +8% over 3 (80 registers) on B200 */
#endif

/* Scattering queue (template parameter QK = entries per thread in shared memory, a power of two; 0 = off).
 * The scattering block costs ~100 warp instructions and, in a medium where a packet crosses a few voxels per
 * scattering event, runs in EVERY warp-iteration with a third of the lanes (whoever ran out of scattering length in
 * the last segment).  With the queue each lane keeps up to QK pre-drawn events {new direction, next scattering
 * length} chained on each other; the block runs only when some lane needs an event and has none queued, and then
 * EVERY lane with a free slot draws one.  Consuming an event is a 16-byte shared-memory load.  The queue is
 * emptied whenever the direction or the medium changes other than by scattering (launch, reflection, refraction,
 * label change): the draws thrown away are independent of everything that happened to the packet, so nothing is
 * biased.  Measured (B200, 1e8 photons, depth 8): cube60 178.9 -> 162.6 ms, cube60b 308.3 -> 286.4, skinvessel
 * 302.2 -> 251.7; but colin27 (3e7) 553 -> 625 and digimouse 367 -> 401, where nearly every segment ends in a
 * scattering event and the block runs every iteration anyway -- engine.cu picks the variant from the mean
 * scattering coefficient per voxel.  Per-thread RNG draw ORDER differs from the reference's, so runs that record
 * seeds for a replay use the kernels without the queue. */
#ifndef MCXB_EXT_MINBLOCKS
    #define MCXB_EXT_MINBLOCKS 4      /* resident blocks per SM of the extended-physics kernels: 64 registers with 68-104 bytes of spills beat 2 blocks
                                         at 87-103 registers by 16-23 % (polarised 99.6 -> 83.5 ms, RF 65.8 -> 53.2, SVMC 89.7 -> 69.3, adjoint 52.9 -> 40.9
                                         for 1e7 photons; profiles/r2_ext_minblocks.log) */
#endif
#ifndef MCXB_UNROLL
    #define MCXB_UNROLL 2
#endif
#ifndef MCXB_LAUNCH_IN_TAIL
    #define MCXB_LAUNCH_IN_TAIL 1
#endif
#ifndef MCXB_GENERIC_SMEM
    #define MCXB_GENERIC_SMEM 0
#endif
constexpr bool kSharedWindow = MCXB_GENERIC_SMEM == 0;
#ifndef MCXB_QUEUE_DEPTH
    #define MCXB_QUEUE_DEPTH 8
#endif
constexpr int kQueueDepth = MCXB_QUEUE_DEPTH;
static_assert(kQueueDepth > 0 && (kQueueDepth & (kQueueDepth - 1)) == 0, "queue depth must be a power of two");

template <int SRC, bool REFLECT, int SAVEDET, typename MediaT, typename AccT, bool STATS, bool GEN, int QDEPTH = 0, bool EXT = false, bool BCODES = false>
__global__ void __launch_bounds__(kBlock, EXT ? MCXB_EXT_MINBLOCKS : MCXB_MINBLOCKS) photon_kernel(const __grid_constant__ SimParam P) {
    extern __shared__ float4 smem[];
    static_assert(QDEPTH == 0 || (!GEN && SAVEDET < 2), "the scattering queue exists for the common-configuration kernels with the default record");
    static_assert(!EXT || GEN, "the extended-physics kernels (polarised light, RF) are generic kernels");
    constexpr uint32_t QK = (uint32_t)QDEPTH;
    float4* const queue = smem + threadIdx.x;             /* entry j of this thread: queue[j * kBlock] */
    float4* tab = smem + QK * kBlock;                     /* optical properties, row 0 = background */
    /* the table through a shared-window address, for the access inside the photon loop (photon_device.cuh) */
    /* Measured (B200, profiles/r2_sweep_shared_window.log): the kernels WITHOUT the scattering queue gain from it (colin27
     * 536.7 -> 522.8 ms, digimouse 355.0 -> 353.3), the queue kernels lose (cube60b 262.5 -> 264.6; with only the table
     * through the window 271.1): used where QDEPTH == 0. */
    constexpr bool useWindow = kSharedWindow && QK == 0;
    uint32_t tab_s = smem_addr(tab);

    if (useWindow) {
        asm volatile("mov.u32 %0, %0;" : "+r"(tab_s));      /* opaque: keeps the window address in ONE register instead of rebuilding it per access */
    }

    const float4* srctab = tab + P.medianum;              /* 4 rows per source, main source first  */
    const float4* dettab = srctab + 4 * (1 + P.extrasrclen);
    float* ftab = reinterpret_cast<float*>(tab + P.tablen);        /* inverse-CDF tables */
    float* ppath_base = ftab + P.ftablen;                         /* partialdata x blockDim, thread-minor */
    unsigned long long* seed_base = reinterpret_cast<unsigned long long*>(ppath_base + (SAVEDET ? P.partialdata * kBlock : 0));

    for (uint32_t i = threadIdx.x; i < P.tablen; i += kBlock) {
        tab[i] = P.tables[i];
    }

    for (uint32_t i = threadIdx.x; i < P.nphase; i += kBlock) {
        ftab[i] = P.invcdf[i];
    }

    for (uint32_t i = threadIdx.x; i < P.nangle; i += kBlock) {
        ftab[P.nphase + i] = P.angleinvcdf[i];
    }

    float* ppath = ppath_base + threadIdx.x;
    unsigned long long* photonseed = seed_base + threadIdx.x;
    /* multi-source runs in the common-configuration kernels: the source a packet came from is needed once per detected
     * photon and, with one volume per source, once per deposit -- kept in shared memory (one word per thread, after the
     * seed rows) so that single-source runs do not carry a live register for it; the generic kernels use a register */
    int* const srcslot = reinterpret_cast<int*>(seed_base + ((SAVEDET && P.issaveseed) ? 2 * kBlock : 0)) + threadIdx.x;

    if (!GEN && P.extrasrclen) {
        *srcslot = 0;
    }

    if (SAVEDET) {
        for (uint32_t i = 0; i < P.partialdata; i++) {
            ppath[i * kBlock] = 0.f;
        }
    }

    __syncthreads();

#ifdef MCXB_EXP_SMEMTILE
    /* measurement only (profiles/r2_deposit_variants.md): a 16^3 fp32 tile of the volume around the source voxel
     * privatised in shared memory per block, flushed with one global reduction per touched voxel at block exit */
    __shared__ float exp_tile[4096];
    int exp_ox, exp_oy, exp_oz;
    {
        const float4 sp = srctab[0];
        exp_ox = max(0, min((int)P.nx - 16, (int)floorf(sp.x) - 8));
        exp_oy = max(0, min((int)P.ny - 16, (int)floorf(sp.y) - 8));
        exp_oz = max(0, min((int)P.nz - 16, (int)floorf(sp.z) - 8));

        for (int i = threadIdx.x; i < 4096; i += kBlock) {
            exp_tile[i] = 0.f;
        }

        __syncthreads();
    }
    int exp_oldtile = -1;
#endif
    const uint32_t tid = blockIdx.x * kBlock + threadIdx.x;
    const MediaT* __restrict__ media = static_cast<const MediaT*>(P.media);
    /* Small volumes are accumulated into several copies (engine.cu picks the count): the few voxels next to the source
     * receive a reduction from nearly every packet, and reductions to one address are serialised by the L2 slice that
     * owns it (measured: the same kernel ran 349 / 361 / 404 ms depending on which physical pages backed a single
     * copy, i.e. on which hot lines happened to share a slice).  Spreading the packets over k copies divides the load
     * on every hot line by k at no cost in the loop: the copy is folded into the base pointer here. */
    const uint32_t copyoff = (uint32_t)((blockIdx.x % P.acccopies) * P.accstride);     /* < 2^32: copies exist for small volumes only */
    AccT* __restrict__ field = static_cast<AccT*>(P.field) + copyoff;
    const float n0 = tab[0].w;
    /* partial-path rows of this thread, biased so that the row of label L is ppath_len[L * kBlock] */
    /* what a detected-photon record holds: read at run time by the generic kernels and by SAVEDET == 2, the default
     * "DP" (detector id + partial paths, src/mcx_utils.c:266) with SAVEDET == 1, where the tests on it then fold away */
    const uint32_t detflag = (GEN || SAVEDET == 2) ? P.savedetflag : 0x5u;
    float* const ppath_len = ppath + ((int)(((detflag >> 1) & 1u) * (P.medianum - 1)) - 1) * kBlock;

    Rng rng;
    rng_seed(rng, P.seeds + 4 * (size_t)tid);

    Photon ph;
    ph.px = ph.py = ph.pz = 0.f;
    ph.w = __int_as_float(0x7FC00000);        /* NaN: nothing to retire before the first launch (:2329) */
    ph.vx = ph.vy = ph.vz = 0.f;
    ph.nscat = -1;
    ph.slen = 0.f;
    ph.tof = 0.f;
    ph.ix = ph.iy = ph.iz = 0;
    ph.face = -1;
    ph.idx1d = 0;
    ph.label = 0;
    ph.detflag = 0;
    ph.w0 = 0.f;
    ph.pathlen = 0.f;
    ph.n1 = n0;

    float mua = 0.f, mus = 0.f, g = 0.f, nmed = n0;   /* optical properties of the voxel being traversed */
    float e_escaped = 0.f, e_launched = 0.f;
    float w0init = 0.f;                               /* launch weight of the live packet (W flag) */
    float pacc = 0.f;                                 /* path length travelled in the current medium, not yet added to its ppath row */
    int   cursrc = 0;                                 /* source the live packet came from (1-based; 0 = single source) */
    uint32_t budget = (P.sched == 1) ? (P.threadphoton + ((int)tid < P.oddphoton ? 1u : 0u)) : 0u;
    uint32_t detarg = 0;                              /* detector argument handed to the retire step */
    uint32_t nextid = (P.sched == 1) ? (tid * P.threadphoton + min(tid, (uint32_t)P.oddphoton)) : 0u;   /* replay: record of the next packet */
    uint32_t curid = 0;                               /* replay: record of the live packet */
    uint32_t patidx = 0;                              /* photon sharing: pattern cell the live packet was launched from */
    bool relaunch = true;
    uint32_t qs = 0;                                  /* scattering queue: events queued (low byte) + 256 x events ever pushed */
    float sI = 1.f, sQ = 0.f, sU = 0.f, sV = 0.f;     /* EXT: Stokes vector of the live packet */
    float wre = 0.f, wim = 0.f, w0re = 0.f, w0im = 0.f;   /* EXT, RF forward: complex weight now and at the last deposit; ph.w is its magnitude */
    float rfcos = 1.f, rfsin = 0.f;                   /* EXT, RF replay: cos / sin of omega x detected time of flight of the replayed record */
    /* EXT, split-voxel media (MED_TYPE 97): the plane of the current voxel as seen from the packet's side, whether the next
     * segment has to be tested against it, and whether the last segment ended on it */
    const bool svmc = EXT && sizeof(MediaT) == 4 && P.mediaformat == 97u;
    SplitVoxel nu = { 0.f, 0.f, 0.f, 0.f, 0u };
    bool testint = true, hitintf = false;
    unsigned long long c_seg = 0, c_dep = 0, c_scat = 0;

    /* Retire the packet that just ended (if any) and launch the next one; returns true when this thread has nothing
     * left to do.  One body, two call sites chosen at compile time (kLaunchInTail below). */
    auto next_packet = [&]() -> bool {
        {
            /* ------------------------------------------------------------------ retire (:1494-1569) */
            if (!(ph.w != ph.w)) {
                e_escaped += ph.w;

                if (GEN && P.trajdata) {      /* where the packet ended (:1497-1503) */
                    save_traj(P, curid + 1u, ph.px, ph.py, ph.pz, ph.w, cursrc);
                }

                if (GEN && P.issaveref == 1 && ph.label == 0 && ph.idx1d != kOutsideMin && ph.idx1d != kOutsideMax && ph.w > 0.f) {
                    int tshift = max(0, min((int)P.maxgate - 1, (int)floorf((ph.tof - P.twin0) * P.Rtstep)));

                    if (P.extrasrclen && P.srcid < 0) {
                        tshift += (cursrc - 1) * (int)(P.nrepvol * P.maxgate);
                    }

                    red_add(field + ph.idx1d + (size_t)tshift * P.dimxyz, -ph.w);
                }

                if (SAVEDET) {
                    if ((detarg & kDetMask) && (ph.label == 0 || (svmc && sv_label(nu.sv) == 0u)) && (!GEN || P.issaveref < 2)) {      /* :1555-1561 */
                        if (EXT) {
                            const float st[4] = { sI, sQ, sU, sV };
                            save_detected(P, dettab, ppath, kBlock, ph, detarg, w0init, cursrc, photonseed, st);
                        } else {
                            save_detected(P, dettab, ppath, kBlock, ph, detarg, w0init, GEN ? cursrc : (P.extrasrclen ? *srcslot : 0), photonseed);
                        }
                    }
                }
            }

            if (SAVEDET) {
                #pragma unroll 1

                for (uint32_t i = 0; i < P.partialdata; i++) {
                    ppath[i * kBlock] = 0.f;
                }
            }

            /* ------------------------------------------------------------------ photon budget */
            if (budget == 0) {
                if (P.sched == 1) {
                    return true;
                }

                /* guided self-scheduling: claim P.chunk photons while plenty are left, fewer as the counter approaches
                 * nphoton (about half of an even share of what remains), so that all threads run dry within one
                 * photon of each other instead of within one chunk.  The peek is a plain load: a stale value only
                 * changes the size of the claim, never its correctness. */
                uint32_t want = P.chunk;
#ifndef MCXB_NO_GUIDED
                {
                    const unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(P.counter);
                    const float left = (seen < P.nphoton) ? (float)(P.nphoton - seen) : 0.f;
                    const float share = left * mufu_rcp(2.f * (float)(gridDim.x * kBlock));
                    want = (uint32_t)fminf((float)P.chunk, fmaxf(1.f, share));
                }
#endif
                const unsigned long long first = atomicAdd(P.counter, (unsigned long long)want);

                if (first >= P.nphoton) {
                    return true;
                }

                budget = (uint32_t)min((unsigned long long)want, P.nphoton - first);

                if (GEN) {
                    nextid = (uint32_t)first;
                }
            }

            /* ------------------------------------------------------------------ replay: restart the stream (:1590-1596) */
            if (GEN && P.replayseed) {
                curid = nextid++;
                rng.a = __ldg(P.replayseed + 2 * (size_t)curid);
                rng.b = __ldg(P.replayseed + 2 * (size_t)curid + 1);
            } else if (GEN && P.trajdata) {
                curid = P.idbase + nextid++;      /* the number of this packet in the whole run */
            }

            /* ------------------------------------------------------------------ launch (:1598-2255) */
            if (SAVEDET && P.issaveseed) {      /* once per packet: a run-time test costs nothing here */
                photonseed[0] = rng.a;
                photonseed[kBlock] = rng.b;
            }

            const float4* S = srctab;

            if (P.extrasrclen && P.srcid != 1) {       /* (:1602-1612) once per packet */
                if (P.srcid > 1) {
                    S = srctab + 4 * (P.srcid - 1);
                } else {
                    const int pick = (int)(rng_uniform(rng) * kJustBelowOne * (float)(P.extrasrclen + 1)) + 1;
                    S = srctab + 4 * (pick - 1);

                    if (GEN) {
                        cursrc = pick;
                    } else {
                        *srcslot = pick;
                    }
                }
            }

            if (GEN && P.replayseed && P.srcid >= 1) {
                (void)rng_uniform(rng);          /* :1614-1616 */
            }

            uint32_t rawlabel = 0, rawdet = 0;
            float attempts = 1.f;
            bool failed = false;

            if (!GEN && SRC == srcPencil && P.plainlaunch) {
                /* the usual pencil beam (engine.cu sets the flag: launch voxel labelled, weight above the roulette
                 * threshold, no focal length, no launch-angle table, one source): nothing is drawn, nothing is tested */
                const float4 pos = srctab[0], dir = srctab[1], p2 = srctab[3];
                ph.px = pos.x;
                ph.py = pos.y;
                ph.pz = pos.z;
                ph.w = pos.w;
                ph.vx = dir.x;
                ph.vy = dir.y;
                ph.vz = dir.z;
                ph.tof = 0.f;
                ph.idx1d = __float_as_uint(p2.z);
                rawlabel = __float_as_uint(p2.w) & 0x7FFFFFFFu;
                rawdet = __float_as_uint(p2.w) & kDetMask;
                ph.ix = (int)(short)floorf(ph.px);
                ph.iy = (int)(short)floorf(ph.py);
                ph.iz = (int)(short)floorf(ph.pz);
            } else do {
                float fx, fy, fz;
                bool aimed;
                ph.slen = 0.f;
                ph.tof = 0.f;
                sample_source<SRC, MediaT>(P, S, rng, ph, rawlabel, rawdet, fx, fy, fz, aimed, patidx);

                if (fabsf(ph.w) <= P.minenergy) {
                    continue;
                }

                const float focal = S[1].w;

                if (P.nangle) {
                    /* launch zenith angle from a user table (:2102-2124) */
                    const float* at = ftab + P.nphase;
                    float ang, c;

                    if (focal > 0.f) {
                        ang = fminf(rng_uniform(rng) * (float)P.nangle, (float)P.nangle - kEps);
                        c = at[(int)ang];
                    } else {
                        ang = fminf(rng_uniform(rng) * (float)(P.nangle - 1), (float)(P.nangle - 1) - kEps);
                        const float fr = ang - (float)(int)ang;
                        const uint32_t i0 = ((uint32_t)ang >= P.nangle - 1) ? P.nangle - 1 : (uint32_t)ang;
                        const uint32_t i1 = ((uint32_t)ang + 1 >= P.nangle - 1) ? P.nangle - 1 : (uint32_t)ang + 1;
                        c = (1.f - fr) * at[i0] + fr * at[i1];
                    }

                    float stheta, ctheta, sphi, cphi;
                    __sincosf(c * kOnePi, &stheta, &ctheta);
                    __sincosf(kTwoPi * rng_uniform(rng), &sphi, &cphi);

                    if (focal < 1.5f && focal >= 0.f) {
                        ph.vx = S[1].x;
                        ph.vy = S[1].y;
                        ph.vz = S[1].z;
                    }

                    rotate_direction(ph.vx, ph.vy, ph.vz, stheta, ctheta, sphi, cphi);
                } else if (aimed) {
                    if (focal != focal) {                       /* NaN: isotropic launch (:2126-2132) */
                        float stheta, ctheta, sphi, cphi;
                        __sincosf(kTwoPi * rng_uniform(rng), &sphi, &cphi);
                        __sincosf(acosf(2.f * rng_uniform(rng) - 1.f), &stheta, &ctheta);
                        rotate_direction(ph.vx, ph.vy, ph.vz, stheta, ctheta, sphi, cphi);
                    } else if (focal < 0.f && isinf(focal)) {   /* -inf: Lambertian launch (:2133-2139) */
                        float sphi, cphi;
                        __sincosf(kTwoPi * rng_uniform(rng), &sphi, &cphi);
                        const float stheta = fast_sqrt(rng_uniform(rng));
                        const float ctheta = fast_sqrt(1.f - stheta * stheta);
                        rotate_direction(ph.vx, ph.vy, ph.vz, stheta, ctheta, sphi, cphi);
                    } else if (focal != 0.f) {                  /* converge to / diverge from a focal point (:2140-2151) */
                        const float sgn = (float)((focal > 0.f) - (focal < 0.f));
                        fx += focal * ph.vx;
                        fy += focal * ph.vy;
                        fz += focal * ph.vz;
                        ph.vx = sgn * (fx - ph.px);
                        ph.vy = sgn * (fy - ph.py);
                        ph.vz = sgn * (fz - ph.pz);
                        const float r = fast_rsqrt(ph.vx * ph.vx + ph.vy * ph.vy + ph.vz * ph.vz);
                        ph.vx *= r;
                        ph.vy *= r;
                        ph.vz *= r;
                    }
                }

                if (EXT && P.adjfirstdet != 0xFFFFFFFFu && (uint32_t)((P.extrasrclen & (P.srcid < 0 ? 1u : 0u)) ? cursrc : 0) > P.adjfirstdet) {
                    /* adjoint runs: a detector is launched as a disk of its radius around its position, perpendicular to
                     * its direction (:2154-2183).  The reference takes the source id from `extrasrclen & (srcid < 0)`, a
                     * bitwise AND: with an even number of extra sources detectors launch as plain points -- kept */
                    float sphi, cphi;
                    __sincosf(kTwoPi * rng_uniform(rng), &sphi, &cphi);
                    const float rad = fast_sqrt(rng_uniform(rng)) * S[2].x;

                    if (ph.vz > -1.f + kEps && ph.vz < 1.f - kEps) {
                        const float t0 = 1.f - ph.vz * ph.vz;
                        const float t1 = rad * fast_rsqrt(t0);
                        ph.px += t1 * (ph.vx * ph.vz * cphi - ph.vy * sphi);
                        ph.py += t1 * (ph.vy * ph.vz * cphi + ph.vx * sphi);
                        ph.pz -= t1 * t0 * cphi;
                    } else {
                        ph.px += rad * cphi;
                        ph.py += rad * sphi;
                    }

                    locate<MediaT>(P, ph, rawlabel, rawdet);
                }

                if (rawlabel == 0) {
                    /* the marcher takes the packet by reference: hand it a copy so that the live packet state never has its
                     * address taken and stays in registers (inlined except in the generic kernels, see enter_volume) */
                    Photon tmp = ph;
                    const int idx = enter_volume<MediaT, GEN && !EXT>(P, tab, tmp);
                    ph = tmp;

                    if (idx >= 0) {
                        ph.idx1d = (uint32_t)idx;
                        fetch_voxel(media, ph.idx1d, rawlabel, rawdet);
                    }
                }

                ph.ix = (int)(short)floorf(ph.px);
                ph.iy = (int)(short)floorf(ph.py);
                ph.iz = (int)(short)floorf(ph.pz);
                attempts += 1.f;

                if (attempts > (float)P.maxvoidstep) {
                    failed = true;
                    break;
                }
            } while (rawlabel == 0 || fabsf(ph.w) <= P.minenergy);

            if (failed) {
                ph.w = __int_as_float(0x7FC00000);
                return true;       /* the source never reaches the volume: this thread gives up (:2204-2206) */
            }

            budget--;
            ph.label = rawlabel;
            ph.detflag = rawdet;

            if (svmc) {          /* :1640-1642, 2217 */
                nu.sv = 0u;
                nmed = svmc_update(P, tab, ph.label, ph.idx1d, ph.px, ph.py, ph.pz, ph.ix, ph.iy, ph.iz, nu).w;
                testint = true;
                hitintf = false;
            } else {
                nmed = medium<MediaT>(P, tab, ph.label).w;
            }

            e_launched += ph.w;
            ph.w0 = ph.w;
            w0init = ph.w;
            /* the first thing the reference does with a fresh packet is draw its scattering length, without a scattering
             * event (the EPS sentinel of :2254 / :2449-2452): done here, so the loop below scatters whenever it draws */
            ph.slen = rng_scatlen(rng);
            ph.nscat = 0;
            ph.pathlen = 0.f;
            ph.face = -1;
            pacc = 0.f;
            qs &= ~0xFFu;        /* queued directions belonged to the previous packet */

            if (EXT) {
                sI = P.s0i;          /* :1633-1638 */
                sQ = P.s0q;
                sU = P.s0u;
                sV = P.s0v;
                wre = w0re = ph.w;   /* :2427-2430 */
                wim = w0im = 0.f;

                if (P.replayseed && (P.outputtype == otRF || P.outputtype == otRFmus)) {      /* :2257-2263 */
                    __sincosf(P.omega * __ldg(P.replaytof + curid), &rfsin, &rfcos);
                }
            }

            if (GEN && P.trajdata) {          /* where the packet starts (:2243-2249) */
                save_traj(P, curid + 1u, ph.px, ph.py, ph.pz, ph.w, cursrc);
            }
        }
        return false;
    };

    /* Where the launch code sits.  The generic kernels test a flag at the top of every iteration.  The common-configuration
     * kernels call it from the tail block, where a packet ends -- no per-iteration test, no loop-carried flag: every
     * thread starts with a DUMMY packet (weight NaN, parked in voxel 0 with label 0, one unit of scattering length left)
     * whose first segment touches nothing (label 0 deposits nothing, a NaN weight is never retired) and whose NaN weight
     * sends it straight into the tail block. */
    constexpr bool kLaunchInTail = !GEN && (MCXB_LAUNCH_IN_TAIL != 0);

    if (kLaunchInTail) {
        ph.px = ph.py = ph.pz = 0.5f;
        ph.vz = 1.f;
        ph.slen = 1.f;
        relaunch = false;
    }

    /* two iterations per trip in the common kernels: ptxas alternates the registers of the loop-carried "previous voxel"
     * values instead of copying them at the top of every iteration (measured -0.3 ... -1.5 % on every deck) */
    constexpr int kUnroll = GEN ? 1 : MCXB_UNROLL;
    #pragma unroll kUnroll

    while (true) {
        if (!kLaunchInTail && relaunch) {
            if (next_packet()) {
                break;
            }

            relaunch = false;
        }

        /* ------------------------------------------------------------------ scattering (:2446-2649) */
        /* (a straight-line version of this block, committed with selects so that ptxas need not rename the packet state
         * around the branch, was measured 1.5 % slower: 319.9 vs 315.2 ms for cube60b 1e8) */
        if (QK > 0) {
            const bool need = ph.slen <= 0.f;

            if (__ballot_sync(__activemask(), need && (qs & 0xFFu) == 0u)) {
                if ((qs & 0xFFu) < QK) {
                    /* one more event for every lane with a free slot, chained on the newest queued direction */
                    float bx = ph.vx, by = ph.vy, bz = ph.vz;

                    if (qs & 0xFFu) {
                        const float4 t = queue[(((qs >> 8) - 1u) & (QK - 1u)) * kBlock];
                        bx = t.x;
                        by = t.y;
                        bz = t.z;
                    }

                    const float gq = tab[ph.label].z;
                    const float ns = rng_scatlen(rng);
                    float sphi, cphi;
                    mufu_sincos(kTwoPi * rng_uniform(rng), sphi, cphi);
                    const float u = rng_uniform(rng);
                    float t = (1.f - gq * gq) * mufu_rcp(1.f - gq + 2.f * gq * u);
                    t *= t;
                    const float chg = fmaxf(-1.f, fminf(1.f, (1.f + gq * gq - t) * mufu_rcp(2.f * gq)));
                    const float ctheta = (fabsf(gq) > kEps) ? chg : (2.f * u - 1.f);
                    const float stheta = fast_sqrt(fmaxf(0.f, 1.f - ctheta * ctheta));
                    rotate_direction(bx, by, bz, stheta, ctheta, sphi, cphi);
                    queue[((qs >> 8) & (QK - 1u)) * kBlock] = make_float4(bx, by, bz, ns);

                    qs += 257u;
                }
            }

            if (need) {
                const float4 t = queue[(((qs >> 8) - (qs & 0xFFu)) & (QK - 1u)) * kBlock];
                ph.vx = t.x;
                ph.vy = t.y;
                ph.vz = t.z;
                ph.slen = t.w;
                qs -= 1u;

                if (SAVEDET && (detflag & 0x02u)) {
                    uint32_t* cnt = reinterpret_cast<uint32_t*>(ppath + (ph.label - 1) * kBlock);
                    *cnt += 1u;
                }

                if (STATS) {
                    c_scat++;
                }
            }
        } else if (ph.slen <= 0.f) {
            ph.slen = rng_scatlen(rng);
            testint = true;          /* SVMC: a new direction may hit the plane again (:2638-2648) */

            {
                float sphi = 0.f, cphi = 1.f, stheta, ctheta;
                const bool flat = GEN && P.is2d;
                /* (a packet that carries label 0 through an in-grid empty voxel has no Mueller matrix: the reference indexes its
                 * table with label - 1 there, :2455; here such a packet scatters by the scalar phase function) */
                const bool polar = EXT && P.maxpolmedia != 0u && !flat && ph.label - 1u < P.maxpolmedia;
                float theta = 0.f, phi = 0.f, s2p = 0.f, c2p = 1.f;

                if (polar) {
                    /* polarised light (:2454-2468): polar angle and azimuth drawn together by rejection against the
                     * intensity the Mueller matrix of this medium scatters into that direction for the current Stokes vector */
                    const float4* sm = P.smatrix + (size_t)kNAngles * (ph.label - 1u);
                    const float4 m0 = __ldg(sm);
                    float i0, i1;

                    do {
                        ctheta = 2.f * rng_uniform(rng) - 1.f;
                        theta = acosf(ctheta);
                        phi = kTwoPi * rng_uniform(rng);
                        __sincosf(phi, &sphi, &cphi);
                        s2p = 2.f * sphi * cphi;            /* sin / cos of 2 phi from the pair the rotation needs anyway */
                        c2p = cphi * cphi - sphi * sphi;
                        const float4 m = __ldg(sm + (uint32_t)(theta * (float)kNAngles * (0.318309886183791f - kEps)));
                        const float qq = sQ * c2p + sU * s2p;
                        i0 = m0.x * sI + m0.y * qq;
                        i1 = m.x * sI + m.y * qq;
                    } while (rng_uniform(rng) * i0 >= i1);

                    stheta = fast_sqrt(fmaxf(0.f, 1.f - ctheta * ctheta));      /* sin(acos(c)) on [0, pi] */
                } else if (!flat) {
                    mufu_sincos(kTwoPi * rng_uniform(rng), sphi, cphi);
                }

                if (polar) {
                    /* angles are set */
                } else if (GEN && P.nphase > 2) {
                    /* tabulated phase function: linear interpolation of the inverse CDF of cos(theta) */
                    float u = rng_uniform(rng) * (float)(P.nphase - 1);
                    const float fr = u - (float)(int)u;
                    const uint32_t i0 = ((uint32_t)u >= P.nphase) ? P.nphase - 1 : (uint32_t)u;
                    const uint32_t i1 = ((uint32_t)u + 1 >= P.nphase) ? P.nphase - 1 : (uint32_t)u + 1;
                    ctheta = (1.f - fr) * ftab[i0] + fr * ftab[i1];
                } else {
                    const float gg = (GEN && (uint32_t)ph.nscat > P.gscatter) ? 0.f : g;

                    /* Henyey-Greenstein inverse CDF (:2487-2490), or the isotropic 2u-1 when |g| <= EPS (:2495-2496);
                     * both from the same draw, chosen with a select (the HG expression is finite garbage for g == 0);
                     * sin(acos(c)) == sqrt(1-c^2) on [0,pi] */
                    const float u = rng_uniform(rng);
                    float t = (1.f - g * g) * mufu_rcp(1.f - g + 2.f * g * u);
                    t *= t;
                    const float chg = fmaxf(-1.f, fminf(1.f, (1.f + g * g - t) * mufu_rcp(2.f * g)));
                    ctheta = (fabsf(gg) > kEps) ? chg : (2.f * u - 1.f);
                }

                if (!polar) {
                    stheta = fast_sqrt(fmaxf(0.f, 1.f - ctheta * ctheta));
                }

                if (SAVEDET) {
                    const uint32_t flag = detflag;
                    const uint32_t M = P.medianum - 1;
                    /* the row of the tissue the event happened in: the label, or the current part of a split voxel (:2503-2531) */
                    const uint32_t row = svmc ? sv_label(nu.sv) : ph.label;

                    if ((flag & 0x02u) && (!svmc || row)) {
                        uint32_t* cnt = reinterpret_cast<uint32_t*>(ppath + (row - 1) * kBlock);
                        *cnt += 1u;
                    }

                    if ((flag & 0x08u) && (!svmc || row)) {
                        ppath[(M * ((flag >> 1 & 1u) + (flag >> 2 & 1u)) + row - 1) * kBlock] += 1.f - ctheta;
                    }
                }

                const float olduz = ph.vz;

                if (flat) {
                    rotate_direction_2d(ph.vx, ph.vy, ph.vz, (rng_uniform(rng) > 0.5f ? stheta : -stheta), ctheta, (int)P.is2d);
                } else {
                    rotate_direction(ph.vx, ph.vy, ph.vz, stheta, ctheta, sphi, cphi);
                }

                ph.nscat++;

                if (polar) {
                    update_stokes(sI, sQ, sU, sV, theta, ctheta, phi, s2p, c2p, olduz, ph.vz, P.smatrix, ph.label);
                }

                if (GEN && P.trajdata) {      /* every scattering site (:2625-2632) */
                    save_traj(P, curid + 1u, ph.px, ph.py, ph.pz, ph.w, cursrc);
                }

                /* scattering-site sensitivities of a replayed packet (:2567-2592): WP counts the events, DCS sums the
                 * momentum transfer 1-cos(theta), WPTOF weights the count by the time of flight; each scaled by the
                 * detected weight and binned by the DETECTED time of flight */
                if (GEN && P.replayseed && (P.outputtype == otWP || P.outputtype == otDCS || P.outputtype == otWPTOF || (EXT && P.outputtype == otRFmus))) {
                    float sw = __ldg(P.replayweight + curid);
                    const float rtof = __ldg(P.replaytof + curid);
                    float swim = 0.f;

                    if (P.outputtype == otDCS) {
                        sw *= 1.f - ctheta;
                    } else if (P.outputtype == otWPTOF) {
                        sw *= rtof;
                    } else if (EXT && P.outputtype == otRFmus) {      /* :2572-2577: detected weight x cos / sin(omega tof) */
                        swim = sw * rfsin;
                        sw *= rfcos;
                    }

                    uint32_t tshift = (uint32_t)max(0, min((int)floorf((rtof - P.twin0) * P.Rtstep), (int)P.maxgate - 1));

                    if (P.replaydet == -1) {
                        tshift += (uint32_t)((__ldg(P.replaydetid + curid) & 0xFFFF) - 1) * P.maxgate;
                    }

                    if (P.extrasrclen && P.srcid < 0) {
                        tshift += (uint32_t)(cursrc - 1) * P.nrepvol * P.maxgate;
                    }

                    red_add(field + ((size_t)tshift * P.dimxyz + ph.idx1d), sw);

                    if (EXT && P.outputtype == otRFmus) {
                        /* imaginary part into the second plane, where the host reads it (src/mcx_host.cpp:1270-1276); the
                         * reference's kernel adds it two planes further (:2601, 2617), outside the buffer its host allocates */
                        red_add(field + (P.rfplane + (size_t)tshift * P.dimxyz + ph.idx1d), swim);
                    }
                }

                if (STATS) {
                    c_scat++;
                }
            }
        }

        /* ------------------------------------------------------------------ one ray segment (:2652-2765) */
        ph.n1 = nmed;
        {
            /* SVMC re-derives the part of the voxel from the position in every iteration (:2666-2669) */
            const float4 pr = svmc ? svmc_update(P, tab, ph.label, ph.idx1d, ph.px, ph.py, ph.pz, ph.ix, ph.iy, ph.iz, nu)
                              : ((useWindow && sizeof(MediaT) < 4) ? lds_f4(tab_s + ph.label * 16u) : medium<MediaT>(P, tab, ph.label));
            mua = pr.x;
            mus = pr.y;
            g = pr.z;
            nmed = pr.w;
        }
        const float dist = face_distance(ph.px, ph.py, ph.pz, ph.vx, ph.vy, ph.vz, ph.ix, ph.iy, ph.iz, ph.face);
        const float musp = (GEN && (uint32_t)(ph.nscat + 1) > P.gscatter) ? __fmul_rn(mus, __fsub_rn(1.f, g)) : mus;
        float slen;
        float len = step_length(dist, musp, ph.slen, slen);

        if (EXT) {
            hitintf = false;

            if (svmc && (nu.sv & 0x10000u) && testint) {
                /* ray_plane_intersect (:1280-1300): the segment ends on the plane if it would cross it */
                const float vdotn = dot3_rn(ph.vx, ph.vy, ph.vz, nu.nx, nu.ny, nu.nz);

                if (vdotn > 0.f) {
                    const float d0 = __fsub_rn(dot3_rn(ph.px, ph.py, ph.pz, nu.nx, nu.ny, nu.nz), nu.pd);
                    const float d1 = __fadd_rn(d0, __fmul_rn(len, vdotn));

                    if (!(__fmul_rn(d0, d1) > 0.f)) {
                        const float len0 = __fdiv_rn(__fmul_rn(len, d0), __fsub_rn(d0, d1));
                        len = (len0 > 0.f) ? len0 : len;
                        slen = __fmul_rn(len, musp);
                        hitintf = true;
                    }
                }
            }
        }

        ph.pathlen += len;
        ph.px = advance(ph.px, len, ph.vx);
        ph.py = advance(ph.py, len, ph.vy);
        ph.pz = advance(ph.pz, len, ph.vz);
        {
            /* reached the face before the scattering site: step into the neighbour across that face */
            const float vf = (ph.face == 0) ? ph.vx : ((ph.face == 1) ? ph.vy : ph.vz);
            const int d = (slen != ph.slen && !(EXT && hitintf)) ? ((vf > 0.f) ? 1 : -1) : 0;
            ph.ix += (ph.face == 0) ? d : 0;
            ph.iy += (ph.face == 1) ? d : 0;
            ph.iz += (ph.face == 2) ? d : 0;
        }
        if (EXT && P.rfforward) {
            /* RF forward (:2750-2763): w <- w exp[-(mua + i omega n / c0) ds]; the magnitude drives roulette and the energy ledger */
            const float att = mufu_ex2(mua * len * -1.4426950408889634f);
            float rs, rc;
            __sincosf(P.omega * nmed * P.oneoverc0 * len, &rs, &rc);
            const float tre = att * (wre * rc + wim * rs), tim = att * (-wre * rs + wim * rc);
            wre = tre;
            wim = tim;
            ph.w = sqrtf(wre * wre + wim * wim);
        } else {
            ph.w *= mufu_ex2(mua * len * -1.4426950408889634f);
        }

        ph.slen -= slen;
        ph.tof += len * nmed * P.oneoverc0;

        if (STATS) {
            c_seg++;
        }

        if (SAVEDET) {
            if (svmc) {      /* split voxels: straight into the row of the current part (:2778-2782) */
                const uint32_t row = sv_label(nu.sv);

                if ((detflag & 0x04u) && row) {
                    ppath_len[row * kBlock] += len;
                }
            } else {
                pacc += len;     /* moved to the partial-path row of this medium when the packet leaves it (below) */
            }
        }

        /* ------------------------------------------------------------------ new voxel (:2796-2811) */
        const uint32_t oldidx = ph.idx1d;
        const uint32_t olddet = ph.detflag;
        const uint32_t oldlabel = ph.label;

        /* the reference compares (ushort)id against the dimension (:2801); ids stay within [-32768, 32767] */
        if ((uint32_t)ph.ix < P.nx && (uint32_t)ph.iy < P.ny && (uint32_t)ph.iz < P.nz) {
            ph.idx1d = linear_index(P, ph.ix, ph.iy, ph.iz);
            fetch_voxel(media, ph.idx1d, ph.label, ph.detflag);
            /* a packet that entered a label-0 voxel INSIDE the grid and was not retired there (no index mismatch, or
             * reflection compiled out) keeps label 0 for the rest of its life: the reference restores
             * mediaid = mediaidold whenever the voxel just left had label 0 (:2816, 2927-2929) */
            ph.label = oldlabel ? ph.label : 0u;
        } else {
            ph.label = 0;
            ph.idx1d = (ph.ix < 0 || ph.iy < 0 || ph.iz < 0) ? kOutsideMin : kOutsideMax;

            if (GEN) {
                const uint32_t code = P.bc[(ph.idx1d == kOutsideMax) * 3 + ph.face];
                ph.detflag = ((code & 0xFu) == bcUnknown) ? (P.doreflect ? (uint32_t)bcReflect : (uint32_t)bcAbsorb) : code;
            } else {
                ph.detflag = REFLECT ? (uint32_t)bcReflect : (uint32_t)bcAbsorb;
            }
        }

        /* ------------------------------------------------------------------ deposit (:2816-2929) */
#ifndef MCXB_NESTED_DEPOSIT
        if (!GEN) {
            /* common configuration (flux / fluence, one volume per gate): ONE divergent region instead of three nested
             * ones -- every level of nesting costs a BSSY / BRA / BSYNC triple per warp-iteration */
            const bool moved = ph.idx1d != oldidx;
            float weight = (mua < kEps) ? (ph.w0 * ph.pathlen) : ((ph.w0 - ph.w) * mufu_rcp(mua));

            if (BCODES) {       /* these kernels also serve the energy and path-length outputs (:2842-2843, 2862-2864), one uniform test */
                weight = (P.outputtype == otEnergy) ? (ph.w0 - ph.w) : ((P.outputtype == otL) ? (ph.w0 * ph.pathlen) : weight);
            }

            /* nothing is deposited into a label-0 voxel (:2816: "&& mediaidold") */
            if (moved && oldlabel && ph.tof >= P.twin0 && ph.tof < P.twin1 && fabsf(weight) > 0.f) {
#if defined(MCXB_EXP_NODEPOSIT)

                if (weight == 123456.789f)
#endif
                {
                    if (P.widedep) {        /* several gates (bit 0) and / or one volume per source (bit 1): ONE uniform test */
                        /* clamped: (tof-twin0)*Rtstep can round up to maxgate for tof one ulp below twin1 */
                        uint32_t gate = (uint32_t)max(0, min((int)floorf((ph.tof - P.twin0) * P.Rtstep), (int)P.maxgate - 1));

                        if (P.widedep & 2u) {
                            gate += (uint32_t)(*srcslot - 1) * P.maxgate;
                        }

                        red_add(static_cast<AccT*>(P.field) + ((size_t)gate * P.dimxyz + (oldidx + copyoff)), weight);
                    } else {
#if defined(MCXB_EXP_MATCHANY)
                        /* measurement only: lanes of this warp that deposit into the SAME voxel in this iteration are summed
                         * first (match.any + shuffles), one reduction per distinct voxel */
                        const uint32_t act = __activemask();
                        const uint32_t peers = __match_any_sync(act, oldidx);
                        const uint32_t lead = __ffs(peers) - 1u;
                        uint32_t rest = peers & ~(1u << lead);
                        float sum = weight;

                        while (__any_sync(act, rest != 0u)) {
                            const int src = rest ? (__ffs(rest) - 1) : (int)lane_id();
                            const float v = __shfl_sync(act, weight, src);

                            if (lane_id() == lead && rest) {
                                sum += v;
                            }

                            rest &= rest - 1u;
                        }

                        if (lane_id() == lead) {
                            red_add(static_cast<AccT*>(P.field) + (oldidx + copyoff), sum);
                        }

#elif defined(MCXB_EXP_SMEMTILE)

                        if (exp_oldtile >= 0) {
                            asm volatile("red.shared.add.f32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(exp_tile + exp_oldtile)), "f"(weight) : "memory");
                        } else {
                            red_add(static_cast<AccT*>(P.field) + (oldidx + copyoff), weight);
                        }

#else
                        red_add(static_cast<AccT*>(P.field) + (oldidx + copyoff), weight);
#endif
                    }
                }

                if (STATS) {
                    c_dep++;
                }
            }

#ifdef MCXB_EXP_SMEMTILE
            {
                const uint32_t tx = (uint32_t)(ph.ix - exp_ox), ty = (uint32_t)(ph.iy - exp_oy), tz = (uint32_t)(ph.iz - exp_oz);
                exp_oldtile = (tx < 16u && ty < 16u && tz < 16u && ph.label) ? (int)(tz * 256u + ty * 16u + tx) : -1;
            }
#endif
            ph.w0 = moved ? ph.w : ph.w0;
            ph.pathlen = moved ? 0.f : ph.pathlen;
        } else
#endif
        {
            /* a segment that ended on the plane of a split voxel deposits like one that ended on a face (:2816-2825) */
            const bool moved = ph.idx1d != oldidx || (EXT && hitintf);

            if (moved && oldlabel && (!GEN || P.save2pt) && ph.tof >= P.twin0 && ph.tof < P.twin1) {
                float weight;
                uint32_t tshift = 0;

                if (GEN && P.maxgate > 1) {
                    /* clamped: (tof-twin0)*Rtstep can round up to maxgate for tof one ulp below twin1 */
                    tshift = (uint32_t)min((int)floorf((ph.tof - P.twin0) * P.Rtstep), (int)P.maxgate - 1);
                }

                float weight_im = 0.f;       /* EXT: what goes into the second (imaginary) plane of an RF output */

                if (EXT && P.rfforward) {
                    /* RF forward (:2833-2841): (w0 - w) / (mua + i omega n / c0), a complex quotient */
                    const float dre = w0re - wre, dim = w0im - wim;
                    const float aim = P.omega * nmed * P.oneoverc0;
                    const float amag2 = mua * mua + aim * aim;
                    weight = (amag2 < kEps) ? (w0re * ph.pathlen) : (dre * mua + dim * aim) / amag2;
                    weight_im = (amag2 < kEps) ? (w0im * ph.pathlen) : (dim * mua - dre * aim) / amag2;
                } else if (GEN && P.outputtype == otEnergy) {
                    weight = ph.w0 - ph.w;
                } else if (GEN && P.outputtype == otL) {
                    weight = ph.w0 * ph.pathlen;
                } else if (GEN && P.outputtype > otEnergy) {
                    /* replay outputs (:2845-2858): the absorption Jacobian deposits detected weight x path length in
                     * the voxel just left (WLTOF: additionally x time of flight), binned by the DETECTED time of
                     * flight and, with replaydet == -1, by detector; WP / DCS / WPTOF deposit at scattering sites */
                    weight = 0.f;

                    if (P.replayseed && (P.outputtype == otJacobian || P.outputtype == otWLTOF || (EXT && P.outputtype == otRF))) {
                        const float rtof = __ldg(P.replaytof + curid);
                        weight = __ldg(P.replayweight + curid) * ph.pathlen * (P.outputtype == otWLTOF ? rtof : 1.f);

                        if (EXT && P.outputtype == otRF) {       /* RF Jacobian (:2853-2855, 2895-2897): -w L {cos, sin}(omega tof) */
                            weight_im = -weight * rfsin;
                            weight = -weight * rfcos;
                        }

                        tshift = (uint32_t)max(0, min((int)floorf((rtof - P.twin0) * P.Rtstep), (int)P.maxgate - 1));

                        if (P.replaydet == -1) {
                            tshift += (uint32_t)((__ldg(P.replaydetid + curid) & 0xFFFF) - 1) * P.maxgate;
                        }
                    }
                } else {
                    weight = (mua < kEps) ? (ph.w0 * ph.pathlen) : ((ph.w0 - ph.w) * mufu_rcp(mua));
                }

                if (GEN && P.extrasrclen && P.srcid < 0) {
                    tshift += (uint32_t)(cursrc - 1) * P.nrepvol * P.maxgate;
                }

                if (GEN && P.srcnum > 1) {
                    /* photon sharing (:2902-2911): volumes interleaved pattern-fastest */
                    const float* pw = P.srcpattern + (size_t)patidx * P.srcnum;
                    AccT* dst = field + ((size_t)tshift * P.dimxyz + oldidx) * P.srcnum;

                    for (uint32_t i = 0; i < P.srcnum; i++) {
                        const float wi = __ldg(pw + i);

                        if (fabsf(wi) > 0.f) {
                            red_add(dst + i, weight * wi);
                        }
                    }
                } else if (fabsf(weight) > 0.f || (EXT && P.rfplane && fabsf(weight_im) > 0.f)) {
                    if (EXT && P.rfplane) {
                        red_add(static_cast<AccT*>(P.field) + (P.rfplane + (size_t)tshift * P.dimxyz + (oldidx + copyoff)), weight_im);
                    }

#if defined(MCXB_EXP_NODEPOSIT)
                    /* timing experiment only (tools/): what the kernel costs without its reductions */
                    if (weight == 123456.789f) {
                        red_add(field + ((size_t)tshift * P.dimxyz + oldidx), weight);
                    }

#elif defined(MCXB_EXP_SPREAD)
                    /* timing experiment only: same number of reductions, hot spots destroyed by a per-thread offset */
                    red_add(field + ((size_t)tshift * P.dimxyz + (oldidx + tid * 977u) % P.dimxyz), weight);
#else
                    /* the usual case -- one gate, one volume -- is a 32-bit index on the (uniform) base pointer; the
                     * common-configuration kernels branch around the gate arithmetic instead of predicating it */
                    if (GEN) {
                        size_t e = (size_t)(oldidx + copyoff);

                        if (tshift) {
                            e += (size_t)tshift * P.dimxyz;
                        }

                        red_add(static_cast<AccT*>(P.field) + e, weight);
                    } else if (P.maxgate > 1) {
                        const uint32_t gate = (uint32_t)min((int)floorf((ph.tof - P.twin0) * P.Rtstep), (int)P.maxgate - 1);
                        red_add(static_cast<AccT*>(P.field) + ((size_t)gate * P.dimxyz + (oldidx + copyoff)), weight);
                    } else {
                        red_add(static_cast<AccT*>(P.field) + (oldidx + copyoff), weight);
                    }
#endif

                    if (STATS) {
                        c_dep++;
                    }
                }
            }

            ph.w0 = moved ? ph.w : ph.w0;
            ph.pathlen = moved ? 0.f : ph.pathlen;

            if (EXT && moved) {
                w0re = wre;
                w0im = wim;
            }
        }

        float4 svprop = make_float4(0.f, 0.f, 0.f, 0.f);      /* SVMC: properties on the far side of the face just crossed */

        if (svmc) {
            /* :2931-2949: a new voxel is looked up at the packet's position; a crossed plane swaps the two parts */
            if (ph.idx1d != oldidx) {
                svprop = svmc_update(P, tab, ph.label, ph.idx1d, ph.px, ph.py, ph.pz, ph.ix, ph.iy, ph.iz, nu);
                testint = true;
            } else if (hitintf) {
                nu.nx = -nu.nx;
                nu.ny = -nu.ny;
                nu.nz = -nu.nz;
                nu.pd = -nu.pd;
                nu.sv ^= 0x20000u;
                testint = false;
            }
        }

        /* Everything below only has work to do for the few packets that changed medium (which includes leaving the
         * grid), ran out of time or fell below the roulette threshold: one test keeps the other lanes out of it.
         * (With the label unchanged n1 == n of the current medium, so the index-mismatch block is a no-op, and
         * boundary codes are only ever attached to a label-0 step out of the grid.) */
        /* (!(|w| >= minenergy) instead of |w| < minenergy: identical for numbers, and true for the NaN weight of the dummy
         * packet every thread starts with when the launch code lives in this block) */
        if (ph.label != oldlabel || ph.tof > P.twin1 || !(fabsf(ph.w) >= P.minenergy) || (svmc && (ph.idx1d != oldidx || hitintf || ph.n1 != nmed))) {
            qs &= ~0xFFu;        /* the medium (g) or the direction may change below: queued events no longer apply */

            if (SAVEDET && !svmc) {
                if (ph.label != oldlabel) {
                    if ((detflag & 0x04u) && oldlabel) {
                        ppath_len[oldlabel * kBlock] += pacc;
                    }

                    pacc = 0.f;
                }
            }

            /* ------------------------------------------------------------------ leave / time out (:2957-3028) */
            /* Per-face boundary codes (`-B`, bc[0..5]) and detect-on-face flags (bc[6..11]) in the common kernels (BCODES
             * instantiations): the loop above attaches the code that isreflect alone implies and the explicit code replaces
             * it here, where a packet that left the grid always arrives -- no instruction in the loop (:2806-2808).  A
             * compile-time axis, not a run-time flag: with the extra tail code present ptxas lays the whole loop out
             * differently, and the decks WITHOUT codes lost up to 23 % (cube60b 261.5 -> 320.9 ms, measured) */
            constexpr bool BC = GEN || BCODES;

            if (!GEN && BCODES && ph.label == 0u && (ph.idx1d == kOutsideMin || ph.idx1d == kOutsideMax)) {
                const uint32_t code = P.bc[(ph.idx1d == kOutsideMax) * 3 + ph.face];
                ph.detflag = ((code & 0xFu) == bcUnknown) ? (P.doreflect ? (uint32_t)bcReflect : (uint32_t)bcAbsorb) : code;
            }

            const uint32_t bcode = ph.detflag & 0xFu;

            /* SVMC (:2962-2966): the packet stepped into the empty lower part of a voxel (through a face or through the plane) */
            const bool svexit = svmc && (ph.idx1d != oldidx || hitintf) && !(nu.sv & 0x20000u) && sv_lower(nu.sv) == 0u && (!P.doreflect || ph.n1 == n0);

            if ((ph.label == 0 && (bcode == bcAbsorb || (BC && bcode == bcCyclic) || (bcode == bcReflect && ph.n1 == n0))) || ph.tof > P.twin1 ||
                    (kLaunchInTail && ph.w != ph.w) || svexit) {
                bool reentered = false;

                if (BC && ph.detflag == bcCyclic) {
                    /* re-enter through the opposite face (:2970-2996) */
                    if (ph.face == 0) {
                        ph.px = nudge(rintf((ph.idx1d == kOutsideMin) ? P.fnx : 0.f), (ph.vx > 0.f) - (ph.vx < 0.f));
                        ph.ix = (int)(short)floorf(ph.px);
                    } else if (ph.face == 1) {
                        ph.py = nudge(rintf((ph.idx1d == kOutsideMin) ? P.fny : 0.f), (ph.vy > 0.f) - (ph.vy < 0.f));
                        ph.iy = (int)(short)floorf(ph.py);
                    } else {
                        ph.pz = nudge(rintf((ph.idx1d == kOutsideMin) ? P.fnz : 0.f), (ph.vz > 0.f) - (ph.vz < 0.f));
                        ph.iz = (int)(short)floorf(ph.pz);
                    }

                    if (voxel_in_grid(P, ph.ix, ph.iy, ph.iz)) {
                        ph.idx1d = linear_index(P, ph.ix, ph.iy, ph.iz);
                        fetch_voxel(media, ph.idx1d, ph.label, ph.detflag);
                        reentered = true;
                    }
                }

                if (!reentered) {
                    detarg = (BC && ((ph.idx1d == kOutsideMax && P.bc[9 + ph.face]) || (ph.idx1d == kOutsideMin && P.bc[6 + ph.face]))) ? kOutsideMin : olddet;
                    relaunch = true;
                }
            } else {
                /* -------------------------------------------------------------- Russian roulette (:3031-3061) */
                if (fabsf(ph.w) < P.minenergy) {
                    if (rng_uniform(rng) * kRouletteSize <= 1.f) {
                        ph.w *= kRouletteSize;

                        if (EXT) {       /* :3035-3040 */
                            wre *= kRouletteSize;
                            wim *= kRouletteSize;
                            w0re *= kRouletteSize;
                            w0im *= kRouletteSize;
                        }
                    } else {
                        detarg = olddet;
                        relaunch = true;
                    }
                }

                /* -------------------------------------------------------------- index mismatch (:3063-3297) */
                if (REFLECT && !relaunch && svmc && hitintf) {
                    /* ---------------------------------------------------------- the plane inside a split voxel (:3074-3111) */
                    const float nlo = tab[sv_lower(nu.sv)].w, nup = tab[sv_upper(nu.sv)].w;

                    if (nlo != nup) {
                        /* reflectray_svmc (:1302-1343), with the normal turned back towards the side the packet came from */
                        nu.nx = -nu.nx;
                        nu.ny = -nu.ny;
                        nu.nz = -nu.nz;
                        nu.pd = -nu.pd;
                        const float icos = fabsf(dot3_rn(ph.vx, ph.vy, ph.vz, nu.nx, nu.ny, nu.nz));
                        const float n2 = (nu.sv & 0x20000u) ? nup : nlo;
                        const float t0 = ph.n1 * ph.n1, t1 = n2 * n2;
                        float t2 = 1.f - t0 / t1 * (1.f - icos * icos);
                        bool reflected = true;

                        if (t2 > 0.f) {
                            float re = t0 * icos * icos + t1 * t2;
                            t2 = sqrtf(t2);
                            const float im = 2.f * ph.n1 * n2 * icos * t2;
                            float rtot = (re - im) / (re + im);
                            re = t1 * icos * icos + t0 * t2 * t2;
                            rtot = (rtot + (re - im) / (re + im)) * 0.5f;
                            reflected = rng_uniform(rng) <= rtot;
                        }

                        if (reflected) {
                            ph.vx += -2.f * icos * nu.nx;
                            ph.vy += -2.f * icos * nu.ny;
                            ph.vz += -2.f * icos * nu.nz;
                            nu.sv ^= 0x20000u;          /* back in the part it came from */
                        } else {
                            const float r = ph.n1 / n2;
                            ph.vx = t2 * nu.nx + r * (ph.vx - icos * nu.nx);
                            ph.vy = t2 * nu.ny + r * (ph.vy - icos * nu.ny);
                            ph.vz = t2 * nu.nz + r * (ph.vz - icos * nu.nz);
                            nu.nx = -nu.nx;
                            nu.ny = -nu.ny;
                            nu.nz = -nu.nz;
                            nu.pd = -nu.pd;

                            if (sv_label(nu.sv) == 0u) {
                                detarg = olddet;       /* refracted into the empty part of the voxel: the packet leaves (:3083-3103) */
                                relaunch = true;
                            } else {
                                nmed = tab[sv_label(nu.sv)].w;
                            }
                        }

                        const float rn = rsqrtf(ph.vx * ph.vx + ph.vy * ph.vy + ph.vz * ph.vz);
                        ph.vx *= rn;
                        ph.vy *= rn;
                        ph.vz *= rn;
                    } else {
                        nmed = tab[sv_label(nu.sv)].w;
                    }
                } else if (REFLECT && !relaunch) {
                    /* SVMC compares with the properties looked up on the far side of the face (:2936-2941, 3143-3147) */
                    const float n2 = svmc ? ((ph.idx1d != oldidx) ? svprop.w : nmed)
                                     : ((ph.label == oldlabel) ? nmed : medium<MediaT>(P, tab, ph.label).w);
                    const bool mirror = BC && bcode == bcMirror;
                    bool handle = false;

                    if (mirror || ph.n1 != n2) {
                        handle = ph.label ? (!BC || P.doreflect)
                                 : (BC ? ((bcode == bcUnknown && P.doreflect) || bcode == bcReflect || bcode == bcMirror)
                                       : (bcode == bcUnknown || bcode == bcReflect));
                    }

                    if (handle) {
                        float Rtotal = 1.f;

                        if (!mirror) {
                            Rtotal = fresnel(ph.vx, ph.vy, ph.vz, ph.n1, n2, ph.face);
                        }

                        if (Rtotal < 1.f && !(ph.label == 0 && mirror) && rng_uniform(rng) > Rtotal) {
                            refract(ph.vx, ph.vy, ph.vz, ph.n1, n2, ph.face);

                            if (ph.label == 0) {
                                detarg = (BC && ((ph.idx1d == kOutsideMax && P.bc[9 + ph.face]) || (ph.idx1d == kOutsideMin && P.bc[6 + ph.face]))) ? kOutsideMin : olddet;
                                relaunch = true;
                            }

                            nmed = n2;     /* now travelling in the new medium */
                        } else {
                            /* mirror the direction and put the packet back on the face it came through (:3204-3213) */
                            if (ph.face == 0) {
                                ph.vx = -ph.vx;
                                ph.px = nudge(rintf(ph.px), 0);
                                ph.ix = (int)(short)rintf(ph.px);
                            } else if (ph.face == 1) {
                                ph.vy = -ph.vy;
                                ph.py = nudge(rintf(ph.py), 0);
                                ph.iy = (int)(short)rintf(ph.py);
                            } else {
                                ph.vz = -ph.vz;
                                ph.pz = nudge(rintf(ph.pz), 0);
                                ph.iz = (int)(short)rintf(ph.pz);
                            }

                            ph.idx1d = oldidx;
                            ph.label = oldlabel;
                            ph.detflag = olddet;
                            nmed = ph.n1;

                            if (svmc) {
                                /* back in the old voxel: which part of it? (:3218-3221); the empty part ends the packet */
                                nmed = svmc_update(P, tab, ph.label, ph.idx1d, ph.px, ph.py, ph.pz, ph.ix, ph.iy, ph.iz, nu).w;

                                if (sv_label(nu.sv) == 0u) {
                                    detarg = olddet;
                                    relaunch = true;
                                }
                            }
                        }
                    } else {
                        nmed = n2;
                    }
                }
            }

            if (kLaunchInTail && relaunch) {
                if (next_packet()) {
                    break;
                }

                relaunch = false;
            }
        }
    }

#ifdef MCXB_EXP_SMEMTILE
    __syncthreads();

    for (int i = threadIdx.x; i < 4096; i += kBlock) {
        const float v = exp_tile[i];

        if (v != 0.f) {
            const uint32_t gi = (uint32_t)(((i >> 8) + exp_oz) * (int)P.dimxy + (((i >> 4) & 15) + exp_oy) * (int)P.nx + ((i & 15) + exp_ox));
            red_add(field + gi, v);
        }
    }

#endif
    /* ---------------------------------------------------------------------- energy bookkeeping (:3301-3302) */
    double esc = (double)e_escaped, lau = (double)e_launched;

    for (int o = 16; o > 0; o >>= 1) {
        esc += __shfl_xor_sync(0xFFFFFFFFu, esc, o);
        lau += __shfl_xor_sync(0xFFFFFFFFu, lau, o);
    }

    if (lane_id() == 0) {
        atomicAdd(P.energy, esc);
        atomicAdd(P.energy + 1, lau);
    }

    if (STATS && P.stats) {
        atomicAdd(P.stats, c_seg);
        atomicAdd(P.stats + 1, c_dep);
        atomicAdd(P.stats + 2, c_scat);
    }
}

} // namespace mcxb
