/*
 * mcx_oracle.c -- plain-C restatement of MCX-CL's photon-transport kernel for label media, split-voxel media (97) and the
 * continuous media formats (Config.mediabyte 99-104), with real or complex (RF forward, omega > 0) packet weights, polarised light and
 * photon replay
 * (TEST INFRASTRUCTURE ONLY: nothing under mcxcl_b200/ may load, link or call this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do).
 *
 * What it restates (reference = fangq/mcxcl, all line numbers are src/mcx_core.cl unless noted):
 *   RNG                        :684-727        xorshift128+, [0,1) floats, scattering-length draw
 *   rotsphi, updatestokes      :792-835        Stokes vector through one scattering event
 *   detector search / records  :838-926
 *   trajectory records         :929-948        -D M / -D T
 *   mcx_nextafterf, hitgrid    :965-995        (OpenCL branch of hitgrid, :988-989)
 *   rotate*, transmit, Fresnel :997-1075
 *   updateproperty             :1079-1193      label rows and the word decoders of MED_TYPE 99-104
 *   split-voxel media          :1231-1344      updateproperty_svmc, ray_plane_intersect, reflectray_svmc
 *   skipvoid                   :1350-1455
 *   launchnewphoton            :1466-2275      all 18 source types, multi-source pick, launch-angle table,
 *                                              focal-length / isotropic / Lambertian launch
 *   mcx_main_loop              :2307-3307      scattering, ray segment, deposit with the MAX_ACCUM shadow
 *                                              spill, termination, cyclic bc, roulette, reflection
 *   host side                  src/mcx_host.cpp:494-524, 674-700, 759-768, 1011-1012 (parameter block, seeding,
 *                              threadphoton/oddphoton), :1252-1306 (fold shadow half, energy sums)
 *   photon replay              :1590-1596, 2568-2612, 2845-2858   stream restart, Jacobian / WP / DCS / WLTOF / WPTOF
 * Not restated: two-word media (96, unreachable from the reference's front-ends), the RF replay outputs (their phase
 * factors are lost in the reference), issaveref > 1 (a data race in the reference).
 *
 * Numeric contract: IEEE binary32, no FMA contraction (build with -ffp-contract=off), the OpenCL native_*
 * functions taken as the libm float functions and rsqrt(x) as 1/sqrtf(x) -- the same contract under which
 * oracle/_ref builds the reference's own kernel source, so that the two can be compared BIT FOR BIT
 * (tests/test_oracle_port.py): this file is pinned by the reference source executed on the same inputs, and
 * through it by the reference's statistical known answers (tests/test_oracle_kat.py).
 *
 * One "work-item" = one RNG stream running threadphoton (+1 for the first oddphoton items) photons in
 * sequence, exactly as one OpenCL work-item does; work-items are independent and are spread over host
 * threads with OpenMP, each host thread owning a private field (summed afterwards).
 */
#include <math.h>
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
    #include <omp.h>
#endif
#include "oracle_api.h"

#define EPS                FLT_EPSILON          /* :476-478 */
#define ONE_PI             3.1415926535897932f
#define TWO_PI             6.28318530717959f
#define R_PI               0.318309886183791f   /* :457 */
#define NANGLES            181                  /* :519 */
#define JUST_BELOW_ONE     0.9998f
#define R_C0               3.335640951981520e-12f
#define ROULETTE_SIZE      10.f                 /* :507 */
#define DET_MASK           0x80000000u          /* :509-514 */
#define MED_MASK           0x7FFFFFFFu
#define MAX_ACCUM          1000.f               /* :516 */
#define OUTSIDE_VOLUME_MIN 0xFFFFFFFFu          /* :494-499 */
#define OUTSIDE_VOLUME_MAX 0x7FFFFFFFu
#define NO_LAUNCH          9999

enum { bcUnknown, bcReflect, bcAbsorb, bcMirror, bcCyclic };
enum { otFlux, otFluence, otEnergy, otJacobian, otWP, otDCS, otRF, otL, otRFmus, otWLTOF, otWPTOF };      /* :666 */

typedef struct { float x, y, z, w; } f4;
typedef struct { short x, y, z, w; } s4;

/* the fields of MCXParam (:611-663) this restatement reads, filled as src/mcx_host.cpp:494-524, 674-694 does */
typedef struct {
    const mcxb_config* cfg;
    f4 maxidx;
    uint32_t dimx, dimxy, dimxyz, fieldlen;      /* dimlen.x/.y/.z/.w */
    float twin0, twin1, oneoverc0, Rtstep, minenergy, minaccumtime;
    uint32_t save2pt, doreflect, savedet, maxdetphoton, maxmedia, detnum;
    int voidtime, srctype, srcid;
    uint32_t maxvoidstep, issaveseed, issaveref, isspecular, maxgate, outputtype;
    uint32_t threadphoton;
    int oddphoton;
    uint32_t debuglevel, savedetflag, reclen, partialdata, w0offset, gscatter, is2d, srcnum, extrasrclen;
    uint32_t nphase, nphaselen, nangle, nanglelen;
    uint32_t maxjumpdebug;                       /* capacity of the trajectory buffer (:929-944) */
    int replay, replaydet;                       /* seed == SEED_FROM_FILE: packets restart from recorded RNG states (:1590-1596) */
    const uint64_t* rseed;
    const float* rweight;
    const float* rtof;
    const int32_t* rdetid;
    uint32_t maxpolmedia;                        /* > 0: polarised run, one Mueller-matrix table per medium (:658) */
    f4 s0;                                       /* incident Stokes vector (:660) */
    const f4* smatrix;                           /* [maxpolmedia][NANGLES] {S11, S12, S33, S43} */
    float omega;                                 /* > 0: RF forward run, complex packet weights (:2427-2430) */
    uint32_t mediaformat;                        /* MED_TYPE of the reference's build: 1 (labels) or 99..104 (:541-546) */
    int doreflection;                            /* MCX_DO_REFLECTION compiled in (src/mcx_host.cpp:945-956) */
    unsigned char bc[12];
    const f4* gproperty;                         /* media rows, then 4 rows per extra source (src/mcx_host.cpp:746-751) */
    const f4* gdetpos;
    const uint32_t* media;
    const float* srcpattern;
    const float* sharedtab;                      /* [nphaselen + nanglelen] inverse-CDF tables (:2354-2383) */
} param_t;

/* buffers private to one host thread */
typedef struct {
    float* field;            /* 2*fieldlen: primary half + shadow half (4*fieldlen in RF forward runs: + imaginary primary / shadow) */
    float* detp;
    uint64_t* detseed;
    uint32_t detcount;
    float* traj;             /* maxjumpdebug x 6 trajectory records of this host thread (-D M / -D T) */
    uint32_t trajcount;
    uint64_t n_segment, n_atomic, n_log;
} sink_t;

/* state of one work-item: the locals of mcx_main_loop (:2331-2349) */
typedef struct {
    f4 p, v, f, prop;
    s4 flipdir;
    uint32_t idx1d, mediaid;
    float w0, Lmove;
    float si, sq, su, sv;    /* Stokes vector of the packet (:603-605, 2341) */
    float svnx, svny, svnz, svpd;   /* split-voxel state MCXsp (:594-598): interface normal, plane offset ... */
    uint32_t svbits;                /* ... and {lower label, upper label, is-split, is-upper} (:577-587) */
    uint64_t t[2];           /* RNG state */
    uint64_t photonseed[2];  /* RNG state at launch (issaveseed) */
    float* ppath;            /* w0offset + srcnum floats (:2396-2399) */
    int threadid;
} item_t;

/* ------------------------------------------------------------------------------------------- RNG */

/* :684-698 */
static float rand_uniform01(uint64_t t[2]) {
    uint64_t s1 = t[0];
    const uint64_t s0 = t[1];
    t[0] = s0;
    s1 ^= s1 << 23;
    t[1] = s1 ^ s0 ^ (s1 >> 18) ^ (s0 >> 5);
    s1 = t[1] + s0;
    union { uint32_t u; float f; } c;
    c.u = 0x3F800000u | ((uint32_t)s1 >> 9);
    return c.f - 1.0f;
}

/* :709-716 */
static void rng_init(uint64_t t[2], const uint32_t* seed) {
    t[0] = (uint64_t)seed[0] << 32 | seed[1];
    t[1] = (uint64_t)seed[2] << 32 | seed[3];
}

static float counted_log(sink_t* s, float x) {
    s->n_log++;
    return logf(x);
}

/* :720-722 */
static float rand_next_scatlen(sink_t* s, uint64_t t[2]) {
    return -counted_log(s, rand_uniform01(t) + EPS);
}

/* ----------------------------------------------------------------------------- small helpers */

static float rsqrt_(float x) {
    return 1.f / sqrtf(x);
}

static uint32_t as_uint(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

/* :178-213 with the host's private buffers: returns the old value */
static float atomicadd(sink_t* s, float* addr, float val) {
    const float old = *addr;
    *addr = old + val;
    s->n_atomic++;
    return old;
}

/* :965-973 */
static float mcx_nextafterf(float a, int dir) {
    union { float f; uint32_t i; } num;
    num.f = a + 1000.f;
    num.i += (uint32_t)dir ^ (num.i & 0x80000000u);
    return num.f - 1000.f;
}

/* :975-995, OpenCL branch: h = |id - (-(v>0)) - p|, h = |(h + EPS) / v|, face = first component equal to the min */
static float hitgrid(sink_t* s, const f4* p0, const f4* v, s4* id) {
    float hx = fabsf(((float)id->x - (v->x > 0.f ? -1.f : 0.f)) - p0->x);
    float hy = fabsf(((float)id->y - (v->y > 0.f ? -1.f : 0.f)) - p0->y);
    float hz = fabsf(((float)id->z - (v->z > 0.f ? -1.f : 0.f)) - p0->z);
    hx = fabsf((hx + EPS) / v->x);
    hy = fabsf((hy + EPS) / v->y);
    hz = fabsf((hz + EPS) / v->z);
    const float dist = fminf(fminf(hx, hy), hz);
    id->w = (short)(dist == hx ? 0 : (dist == hy ? 1 : 2));

    if (s) {
        s->n_segment++;
    }

    return dist;
}

/* :997-1006 */
static void rotate_perpendicular_vector(f4* d, float ax, float ay, float az, float stheta, float ctheta) {
    const float cx = ay * d->z - az * d->y;
    const float cy = az * d->x - ax * d->z;
    const float cz = ax * d->y - ay * d->x;
    d->x = d->x * ctheta + cx * stheta;
    d->y = d->y * ctheta + cy * stheta;
    d->z = d->z * ctheta + cz * stheta;
}

/* :1008-1023 */
static void rotatevector2d(f4* v, float stheta, float ctheta, int is2d) {
    f4 n = *v;

    if (is2d == 1) {
        n.x = 0.f;
        n.y = v->y * ctheta - v->z * stheta;
        n.z = v->y * stheta + v->z * ctheta;
    } else if (is2d == 2) {
        n.x = v->x * ctheta - v->z * stheta;
        n.y = 0.f;
        n.z = v->x * stheta + v->z * ctheta;
    } else if (is2d == 3) {
        n.x = v->x * ctheta - v->y * stheta;
        n.y = v->x * stheta + v->y * ctheta;
        n.z = 0.f;
    }

    *v = n;
    const float tmp0 = rsqrt_(v->x * v->x + v->y * v->y + v->z * v->z);
    v->x *= tmp0;
    v->y *= tmp0;
    v->z *= tmp0;
}

/* :1025-1042 -- note the renormalisation: x, then y, then z, each with the components already updated */
static void rotatevector(f4* v, float stheta, float ctheta, float sphi, float cphi) {
    if (v->z > -1.f + EPS && v->z < 1.f - EPS) {
        const float tmp0 = 1.f - v->z * v->z;
        const float tmp1 = stheta * rsqrt_(tmp0);
        const float nx = tmp1 * (v->x * v->z * cphi - v->y * sphi) + v->x * ctheta;
        const float ny = tmp1 * (v->y * v->z * cphi + v->x * sphi) + v->y * ctheta;
        const float nz = -tmp1 * tmp0 * cphi + v->z * ctheta;
        v->x = nx;
        v->y = ny;
        v->z = nz;
    } else {
        const float nz = (v->z > 0.f) ? ctheta : -ctheta;
        v->x = stheta * cphi;
        v->y = stheta * sphi;
        v->z = nz;
    }

    v->x *= rsqrt_(v->x * v->x + v->y * v->y + v->z * v->z);
    v->y *= rsqrt_(v->x * v->x + v->y * v->y + v->z * v->z);
    v->z *= rsqrt_(v->x * v->x + v->y * v->y + v->z * v->z);
}

/* :1044-1055 */
static void transmit(f4* v, float n1, float n2, short flipdir) {
    const float tmp0 = n1 / n2;
    v->x *= tmp0;
    v->y *= tmp0;
    v->z *= tmp0;

    if (flipdir == 0) {
        v->x = sqrtf(1.f - v->y * v->y - v->z * v->z) * (float)((v->x > 0.f) - (v->x < 0.f));
    } else if (flipdir == 1) {
        v->y = sqrtf(1.f - v->x * v->x - v->z * v->z) * (float)((v->y > 0.f) - (v->y < 0.f));
    } else {
        v->z = sqrtf(1.f - v->x * v->x - v->y * v->y) * (float)((v->z > 0.f) - (v->z < 0.f));
    }
}

/* :1057-1075 */
static float reflectcoeff(const f4* v, float n1, float n2, short flipdir) {
    const float Icos = fabsf((flipdir == 0) ? v->x : (flipdir == 1 ? v->y : v->z));
    const float tmp0 = n1 * n1;
    const float tmp1 = n2 * n2;
    float tmp2 = 1.f - tmp0 / tmp1 * (1.f - Icos * Icos);

    if (tmp2 > 0.f) {
        float Re, Im, Rtotal;
        Re = tmp0 * Icos * Icos + tmp1 * tmp2;
        tmp2 = sqrtf(tmp2);
        Im = 2.f * n1 * n2 * Icos * tmp2;
        Rtotal = (Re - Im) / (Re + Im);
        Re = tmp1 * Icos * Icos + tmp0 * tmp2 * tmp2;
        Rtotal = (Rtotal + (Re - Im) / (Re + Im)) * 0.5f;
        return Rtotal;
    }

    return 1.f;
}

static int in_grid(const param_t* g, const s4* id) {
    /* the reference compares (ushort)id with the float dimension (:1360, 2801) */
    return (float)(unsigned short)id->x < g->maxidx.x && (float)(unsigned short)id->y < g->maxidx.y && (float)(unsigned short)id->z < g->maxidx.z;
}

static uint32_t voxel_index(const param_t* g, const s4* id) {
    return (uint32_t)id->z * g->dimxy + (uint32_t)id->y * g->dimx + (uint32_t)id->x;
}

static int outside_f(const param_t* g, const f4* p) {
    return p->x < 0.f || p->y < 0.f || p->z < 0.f || p->x >= g->maxidx.x || p->y >= g->maxidx.y || p->z >= g->maxidx.z;
}

static void set_voxel_from_pos(s4* id, const f4* p) {
    id->x = (short)floorf(p->x);
    id->y = (short)floorf(p->y);
    id->z = (short)floorf(p->z);
}

/* ------------------------------------------------------------------------- detection :838-926 */

static uint32_t finddetector(const param_t* g, const f4* p0) {
    for (uint32_t i = 0; i < g->detnum; i++) {
        const f4 d = g->gdetpos[i];

        if ((d.x - p0->x) * (d.x - p0->x) + (d.y - p0->y) * (d.y - p0->y) + (d.z - p0->z) * (d.z - p0->z) < d.w * d.w) {
            return i + 1;
        }
    }

    return 0;
}

static void savedetphoton(const param_t* g, sink_t* s, const item_t* it, const float* ppath, uint32_t isdet) {
    int detid = (isdet == OUTSIDE_VOLUME_MIN) ? -1 : (int)finddetector(g, &it->p);

    if (!detid) {
        return;
    }

    uint32_t baseaddr = s->detcount++;

    if (baseaddr >= g->maxdetphoton) {
        return;      /* counted, not stored (the host warns, src/mcx_host.cpp:1207-1210) */
    }

    if (g->issaveseed && s->detseed) {
        s->detseed[2 * (size_t)baseaddr] = it->photonseed[0];
        s->detseed[2 * (size_t)baseaddr + 1] = it->photonseed[1];
    }

    float* rec = s->detp + (size_t)baseaddr * g->reclen;
    const uint32_t flag = g->savedetflag;

    if (flag & 0x01u) {
        if (g->extrasrclen * (g->srcid <= 0)) {
            detid |= ((int)ppath[g->w0offset - 1]) << 16;
        }

        *rec++ = (float)detid;
    }

    for (uint32_t i = 0; i < g->partialdata; i++) {
        *rec++ = ppath[i];
    }

    if (flag & 0x10u) {
        *rec++ = it->p.x;
        *rec++ = it->p.y;
        *rec++ = it->p.z;
    }

    if (flag & 0x20u) {
        *rec++ = it->v.x;
        *rec++ = it->v.y;
        *rec++ = it->v.z;
    }

    if (flag & 0x40u) {
        *rec++ = ppath[g->w0offset - 2];
    }

    if (flag & 0x80u) {      /* :917-922 */
        *rec++ = it->si;
        *rec++ = it->sq;
        *rec++ = it->su;
        *rec++ = it->sv;
    }
}

/* one fluence deposit with the reference's accumulation-precision guard (:2882-2887) */
static void deposit(const param_t* g, sink_t* s, size_t at, float weight) {
    const float oldval = atomicadd(s, s->field + at, weight);

    if (fabsf(oldval) > MAX_ACCUM) {
        atomicadd(s, s->field + at, (oldval > 0.f) ? -MAX_ACCUM : MAX_ACCUM);
        atomicadd(s, s->field + at + g->fieldlen, (oldval > 0.f) ? MAX_ACCUM : -MAX_ACCUM);
    }
}

/* --------------------------------------------------------------------------- skipvoid :1350-1455 */

static int skipvoid(const param_t* g, sink_t* s, f4* p, f4* v, f4* f, s4* flipdir) {
    int count = 1, idx1d;
    set_voxel_from_pos(flipdir, p);
    flipdir->w = -1;

    while (1) {
        if (in_grid(g, flipdir)) {
            idx1d = (int)voxel_index(g, flipdir);

            if (g->media[idx1d] & MED_MASK) {
                p->x -= v->x;
                p->y -= v->y;
                p->z -= v->z;
                set_voxel_from_pos(flipdir, p);
                f->y -= g->minaccumtime;
                idx1d = (int)voxel_index(g, flipdir);
                count = 0;

                while (!in_grid(g, flipdir) || !(g->media[idx1d] & MED_MASK)) {
                    const float dist = hitgrid(s, p, v, flipdir);
                    f->y += g->minaccumtime * dist;
                    p->x = p->x + dist * v->x;
                    p->y = p->y + dist * v->y;
                    p->z = p->z + dist * v->z;

                    if (flipdir->w == 0) {
                        flipdir->x += (v->x > 0.f ? 1 : -1);
                    }

                    if (flipdir->w == 1) {
                        flipdir->y += (v->y > 0.f ? 1 : -1);
                    }

                    if (flipdir->w == 2) {
                        flipdir->z += (v->z > 0.f ? 1 : -1);
                    }

                    idx1d = (int)voxel_index(g, flipdir);

                    if (count++ > 3) {
                        break;
                    }
                }

                f->y = g->voidtime ? f->y : 0.f;

                /* the reference reads media[idx1d] here without a bounds check (:1420-1429); after a failed
                 * refinement the index may lie outside the grid, where this restatement reads label 0 */
                if (g->isspecular) {      /* label media only: the word itself is the table index here (:1424-1431) */
                    const uint32_t lab = ((uint32_t)idx1d < g->dimxyz) ? (g->media[idx1d] & MED_MASK) : 0u;
                    const float nin = g->gproperty[lab].w;

                    if (nin != g->gproperty[0].w) {
                        p->w *= 1.f - reflectcoeff(v, g->gproperty[0].w, nin, flipdir->w);

                        if (p->w > EPS) {
                            transmit(v, g->gproperty[0].w, nin, flipdir->w);
                        }
                    }
                }

                return idx1d;
            }
        }

        if ((p->x < 0.f && v->x <= 0.f) || (p->x >= g->maxidx.x && v->x >= 0.f)
                || (p->y < 0.f && v->y <= 0.f) || (p->y >= g->maxidx.y && v->y >= 0.f)
                || (p->z < 0.f && v->z <= 0.f) || (p->z >= g->maxidx.z && v->z >= 0.f)) {
            return -1;
        }

        p->x = p->x + v->x;
        p->y = p->y + v->y;
        p->z = p->z + v->z;
        set_voxel_from_pos(flipdir, p);
        f->y += g->minaccumtime;

        if ((uint32_t)count++ > g->maxvoidstep) {
            return -1;
        }
    }
}

/* -------------------------------------------------------------------- launchnewphoton :1466-2275 */

/* launch-time media lookup shared by the area sources (:1752-1758) */
#define SV_LOWER(sv)       ((uint32_t)((sv) & 0xFFu))
#define SV_UPPER(sv)       ((uint32_t)(((sv) >> 8) & 0xFFu))
#define SV_ISSPLIT(sv)     (((sv) >> 16) & 1u)
#define SV_ISUPPER(sv)     (((sv) >> 17) & 1u)
#define SV_CURLABEL(sv)    (SV_ISUPPER(sv) ? SV_UPPER(sv) : SV_LOWER(sv))

static float dot3(float ax, float ay, float az, float bx, float by, float bz) {      /* :961-963 */
    return ax * bx + ay * by + az * bz;
}

/* updateproperty_svmc (:1231-1279): split-voxel media, two words per voxel -- {lower, upper, px, py} in the first,
 * {pz, nx, ny, nz} in the second (dimxyz words further on).  Decides from the packet's position which side of the
 * interface it is on and orients the normal towards the other side */
static void update_property_svmc(const param_t* g, item_t* it, f4* prop, uint32_t mediaid, uint32_t idx1d, const f4* p) {
    if (idx1d == OUTSIDE_VOLUME_MIN || idx1d == OUTSIDE_VOLUME_MAX) {
        *prop = g->gproperty[0];
        return;
    }

    const uint32_t w0 = g->media[idx1d + g->dimxyz], w1 = mediaid & MED_MASK;
    const uint32_t c0 = w0 & 0xFFu, c1 = (w0 >> 8) & 0xFFu, c2 = (w0 >> 16) & 0xFFu, c3 = w0 >> 24;
    const uint32_t c4 = w1 & 0xFFu, c5 = (w1 >> 8) & 0xFFu, c6 = (w1 >> 16) & 0xFFu, c7 = w1 >> 24;
    uint32_t svpacked = c7 | (c6 << 8);

    if (c6) {
        const float rx = (float)c5 * (1.f / 255.f) + (float)it->flipdir.x;
        const float ry = (float)c4 * (1.f / 255.f) + (float)it->flipdir.y;
        const float rz = (float)c3 * (1.f / 255.f) + (float)it->flipdir.z;
        it->svnx = (float)c2 * (2.f / 255.f) - 1.f;
        it->svny = (float)c1 * (2.f / 255.f) - 1.f;
        it->svnz = (float)c0 * (2.f / 255.f) - 1.f;
        const float r = rsqrt_(dot3(it->svnx, it->svny, it->svnz, it->svnx, it->svny, it->svnz));
        it->svnx = it->svnx * r;
        it->svny = it->svny * r;
        it->svnz = it->svnz * r;
        it->svpd = dot3(rx, ry, rz, it->svnx, it->svny, it->svnz);

        if (dot3(p->x, p->y, p->z, it->svnx, it->svny, it->svnz) > it->svpd) {
            *prop = g->gproperty[SV_UPPER(svpacked)];
            svpacked |= 0x20000u;
            it->svnx = -it->svnx;
            it->svny = -it->svny;
            it->svnz = -it->svnz;
            it->svpd = -it->svpd;
        } else {
            *prop = g->gproperty[SV_LOWER(svpacked)];
            svpacked &= ~0x20000u;
        }

        svpacked |= 0x10000u;
    } else {
        *prop = g->gproperty[c7];
        svpacked &= 0xFFFFu;
    }

    it->svbits = svpacked;
}

/* ray_plane_intersect (:1281-1302): does the segment of length *len reach the interface?  If so it ends there */
static int ray_plane_intersect(const param_t* g, const item_t* it, const f4* prop, float* len, float* slen) {
    const f4* p0 = &it->p;
    const f4* v = &it->v;
    const float vdotn = dot3(v->x, v->y, v->z, it->svnx, it->svny, it->svnz);

    if (vdotn <= 0.f) {
        return 0;
    }

    const float d0 = dot3(p0->x, p0->y, p0->z, it->svnx, it->svny, it->svnz) - it->svpd;
    const float d1 = d0 + (*len) * vdotn;

    if (d0 * d1 > 0.f) {
        return 0;
    }

    const float len0 = ((*len) * d0) / (d0 - d1);
    *len = (len0 > 0.f) ? len0 : *len;
    *slen = (*len) * prop->y * (v->w + 1.f > (float)g->gscatter ? (1.f - prop->z) : 1.f);
    return 1;
}

/* reflectray_svmc (:1304-1344): Fresnel reflection or refraction at the interface inside a split voxel; returns 1 when the
 * packet is transmitted into an empty (label 0) part */
static int reflectray_svmc(const param_t* g, item_t* it, float n1, float c0[3], f4* prop) {
    float Re, Im, Rtotal, tmp0, tmp1, tmp2;
    const float Icos = fabsf(dot3(c0[0], c0[1], c0[2], it->svnx, it->svny, it->svnz));
    const float n2 = SV_ISUPPER(it->svbits) ? g->gproperty[SV_UPPER(it->svbits)].w : g->gproperty[SV_LOWER(it->svbits)].w;
    tmp0 = n1 * n1;
    tmp1 = n2 * n2;
    tmp2 = 1.f - tmp0 / tmp1 * (1.f - Icos * Icos);

    if (tmp2 > 0.f) {
        Re = tmp0 * Icos * Icos + tmp1 * tmp2;
        tmp2 = sqrtf(tmp2);
        Im = 2.f * n1 * n2 * Icos * tmp2;
        Rtotal = (Re - Im) / (Re + Im);
        Re = tmp1 * Icos * Icos + tmp0 * tmp2 * tmp2;
        Rtotal = (Rtotal + (Re - Im) / (Re + Im)) * 0.5f;

        if (rand_uniform01(it->t) <= Rtotal) {
            c0[0] += (-2.f * Icos) * it->svnx;
            c0[1] += (-2.f * Icos) * it->svny;
            c0[2] += (-2.f * Icos) * it->svnz;
            it->svbits ^= 0x20000u;
        } else {
            c0[0] += (-Icos) * it->svnx;
            c0[1] += (-Icos) * it->svny;
            c0[2] += (-Icos) * it->svnz;
            c0[0] = tmp2 * it->svnx + (n1 / n2) * c0[0];
            c0[1] = tmp2 * it->svny + (n1 / n2) * c0[1];
            c0[2] = tmp2 * it->svnz + (n1 / n2) * c0[2];
            it->svnx = -it->svnx;
            it->svny = -it->svny;
            it->svnz = -it->svnz;
            it->svpd = -it->svpd;

            if (SV_CURLABEL(it->svbits) == 0) {
                return 1;
            }

            *prop = g->gproperty[SV_CURLABEL(it->svbits)];
        }
    } else {
        c0[0] += (-2.f * Icos) * it->svnx;
        c0[1] += (-2.f * Icos) * it->svny;
        c0[2] += (-2.f * Icos) * it->svnz;
        it->svbits ^= 0x20000u;
    }

    tmp0 = rsqrt_(dot3(c0[0], c0[1], c0[2], c0[0], c0[1], c0[2]));
    c0[0] = c0[0] * tmp0;
    c0[1] = c0[1] * tmp0;
    c0[2] = c0[2] * tmp0;
    return 0;
}

/* binary16 storage bits widened exactly to binary32 (what vload_half / convert_float of cl_khr_fp16 do) */
static float half_bits_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1Fu, m = h & 0x3FFu, out;

    if (e == 0x1Fu) {
        out = sign | 0x7F800000u | (m << 13);
    } else if (e) {
        out = sign | ((e + 112u) << 23) | (m << 13);
    } else if (m) {
        e = 113u;

        while (!(m & 0x400u)) {
            m <<= 1;
            e--;
        }

        out = sign | (e << 23) | ((m & 0x3FFu) << 13);
    } else {
        out = sign;
    }

    float f;
    memcpy(&f, &out, 4);
    return f;
}

/* updateproperty (:1079-1193): the optical properties of a voxel from its media word.  Label volumes read a row of the
 * table; the continuous formats decode the word and leave the members they do not carry as they are (prop starts as
 * row 1 at every launch, :2213) */
static void update_property(const param_t* g, f4* prop, uint32_t mediaid) {
    const uint32_t w = mediaid & MED_MASK;
    float mf;

    switch (g->mediaformat) {
        case 101:      /* MEDIA_MUA_FLOAT (:1093-1096) */
            memcpy(&mf, &mediaid, 4);
            prop->x = fabsf(mf);
            prop->w = g->gproperty[w != 0].w;
            break;

        case 100:      /* MEDIA_AS_F2H */
        case 102:      /* MEDIA_AS_HALF (:1103-1123) */
            prop->x = fabsf(half_bits_to_float((uint16_t)(w & 0xFFFFu)));
            prop->y = fabsf(half_bits_to_float((uint16_t)(w >> 16)));
            prop->w = g->gproperty[w != 0].w;
            break;

        case 99: {     /* MEDIA_LABEL_HALF (:1130-1151): a label row with one member replaced */
            const uint32_t lo = w & 0xFFFFu;
            float member[4];
            *prop = g->gproperty[lo & 0x3FFFu];
            memcpy(member, prop, sizeof(member));
            member[(lo & 0xC000u) >> 14] = fabsf(half_bits_to_float((uint16_t)(w >> 16)));
            memcpy(prop, member, sizeof(member));
            break;
        }

        case 103: {    /* MEDIA_ASGN_BYTE (:1157-1169): bytes scale between rows 1 and 2 */
            const f4 lo = g->gproperty[1], hi = g->gproperty[2];
            prop->x = (float)(w & 0xFFu) * (1.f / 255.f) * (hi.x - lo.x) + lo.x;
            prop->y = (float)((w >> 8) & 0xFFu) * (1.f / 255.f) * (hi.y - lo.y) + lo.y;
            prop->z = (float)((w >> 16) & 0xFFu) * (1.f / 255.f) * (hi.z - lo.z) + lo.z;
            prop->w = (float)((w >> 24) & 0xFFu) * (1.f / 127.f) * (hi.w - lo.w) + lo.w;
            break;
        }

        case 104: {    /* MEDIA_AS_SHORT (:1176-1186) */
            const f4 lo = g->gproperty[1], hi = g->gproperty[2];
            prop->x = (float)(w & 0xFFFFu) * (1.f / 65535.f) * (hi.x - lo.x) + lo.x;
            prop->y = (float)(w >> 16) * (1.f / 65535.f) * (hi.y - lo.y) + lo.y;
            prop->w = g->gproperty[w != 0].w;
            break;
        }

        default:       /* label volumes (:1086) */
            *prop = g->gproperty[w];
    }
}

/* rotsphi + updatestokes (:792-835): rotate the Stokes vector into the scattering plane, apply the Mueller matrix of the
 * medium at the scattering angle, rotate back into the new meridian plane, renormalise to I = 1 */
static void updatestokes(const param_t* g, item_t* it, float theta, float phi, const f4* u, const f4* u2) {
    const float costheta = cosf(theta);
    const float sin2phi = sinf(2.f * phi), cos2phi = cosf(2.f * phi);
    float i2 = it->si, q2 = it->sq * cos2phi + it->su * sin2phi, u2s = -it->sq * sin2phi + it->su * cos2phi, v2 = it->sv;
    const uint32_t imedia = NANGLES * ((it->mediaid & MED_MASK) - 1);
    const uint32_t ithedeg = (uint32_t)(theta * NANGLES * (R_PI - EPS));
    const f4 m = g->smatrix[imedia + ithedeg];
    it->si = m.x * i2 + m.y * q2;
    it->sq = m.y * i2 + m.x * q2;
    it->su = m.z * u2s + m.w * v2;
    it->sv = -m.w * u2s + m.z * v2;
    float temp = (u2->z > -1.f && u2->z < 1.f) ? rsqrt_((1.f - costheta * costheta) * (1.f - u2->z * u2->z)) : 0.f;
    float cosi = (temp == 0.f) ? 0.f : (((phi > ONE_PI && phi < TWO_PI) ? 1.f : -1.f) * (u2->z * costheta - u->z) * temp);
    cosi = fmaxf(-1.f, fminf(cosi, 1.f));
    const float sini = sqrtf(1.f - cosi * cosi);
    const float cos22 = 2.f * cosi * cosi - 1.f;
    const float sin22 = 2.f * sini * cosi;
    i2 = it->si;
    q2 = it->sq * cos22 - it->su * sin22;
    u2s = it->sq * sin22 + it->su * cos22;
    v2 = it->sv;
    temp = 1.f / i2;
    it->sq = q2 * temp;
    it->su = u2s * temp;
    it->sv = v2 * temp;
    it->si = 1.f;
}

/* savedebugdata (:929-948): one trajectory record {packet id, x, y, z, weight, source id} through one counter */
static void savedebugdata(const param_t* g, sink_t* s, const f4* p, uint32_t id, int srcid) {
    const uint32_t pos = s->trajcount++;

    if (pos < g->maxjumpdebug && s->traj) {
        float* rec = s->traj + (size_t)pos * 6;
        memcpy(rec, &id, 4);
        rec[1] = p->x;
        rec[2] = p->y;
        rec[3] = p->z;
        rec[4] = p->w;
        rec[5] = (float)srcid;
    }
}

static void locate(const param_t* g, item_t* it) {
    it->idx1d = (uint32_t)((int)floorf(it->p.z)) * g->dimxy + (uint32_t)((int)floorf(it->p.y)) * g->dimx + (uint32_t)((int)floorf(it->p.x));
    it->mediaid = outside_f(g, &it->p) ? 0u : g->media[it->idx1d];
}

/* returns 0 when a packet was launched, non-zero when this work-item is finished */
static int launchnewphoton(const param_t* g, sink_t* s, item_t* it, uint32_t isdet) {
    f4* p = &it->p;
    f4* v = &it->v;
    f4* f = &it->f;
    f4* prop = &it->prop;
    float* ppath = it->ppath;
    uint64_t* t = it->t;
    const mcxb_source* src = &g->cfg->src;
    it->w0 = 1.f;
    it->Lmove = -1.f;

    /* retire the current packet (:1494-1569) */
    if (fabsf(p->w) >= 0.f) {
        ppath[g->partialdata] += p->w;

        if (g->debuglevel & (2u | 8u)) {       /* the end of a trajectory (:1497-1503) */
            savedebugdata(g, s, p, (uint32_t)f->w + (uint32_t)it->threadid * g->threadphoton + (uint32_t)(it->threadid < g->oddphoton ? it->threadid : g->oddphoton),
                          (int)ppath[g->w0offset - 1]);
        }

        if (it->mediaid == 0 && it->idx1d != OUTSIDE_VOLUME_MIN && it->idx1d != OUTSIDE_VOLUME_MAX && g->issaveref && p->w > 0.f) {
            if (g->issaveref == 1) {
                int tshift = (int)g->maxgate - 1;
                const int tg = (int)floorf((f->y - g->twin0) * g->Rtstep);

                if (tg < tshift) {
                    tshift = tg;
                }

                if (g->extrasrclen * (g->srcid < 0)) {
                    tshift += ((int)ppath[g->w0offset - 1] - 1) * (int)g->maxgate;
                }

                if (g->srctype != MCXB_SRC_PATTERN && g->srctype != MCXB_SRC_PATTERN3D) {
                    deposit(g, s, (size_t)(it->idx1d + (uint32_t)tshift * g->dimxyz), -p->w);
                } else {
                    for (uint32_t i = 0; i < g->srcnum; i++) {
                        if (fabsf(ppath[g->w0offset + i]) > 0.f) {
                            deposit(g, s, (size_t)(it->idx1d + (uint32_t)tshift * g->dimxyz) * g->srcnum + i,
                                    -((g->srcnum == 1) ? p->w : p->w * ppath[g->w0offset + i]));
                        }
                    }
                }
            }
        }

        if (g->savedet && (isdet & DET_MASK) == DET_MASK && (it->mediaid == 0 || (g->mediaformat == 97 && SV_CURLABEL(it->svbits) == 0)) && g->issaveref < 2) {      /* :1556-1561 */
            savedetphoton(g, s, it, ppath, isdet);
        }
    }

    if (g->savedet) {
        for (uint32_t i = 0; i < g->partialdata; i++) {
            ppath[i] = 0.f;                                  /* clearpath, :1572 */
        }
    }

    if (f->w >= (float)(g->threadphoton + (it->threadid < g->oddphoton))) {
        return 1;                                            /* :1581 */
    }

    if (g->replay) {      /* :1590-1596, the record index exactly as written */
        const int rec = it->threadid * (int)g->threadphoton + (it->threadid < g->oddphoton - 1 ? it->threadid : g->oddphoton - 1) + ((int)f->w > 0 ? (int)f->w : 0);
        t[0] = g->rseed[2 * (size_t)rec];
        t[1] = g->rseed[2 * (size_t)rec + 1];
    }

    if (g->issaveseed) {
        it->photonseed[0] = t[0];
        it->photonseed[1] = t[1];
    }

    /* multi-source pick (:1602-1612): extra sources live after the media table */
    if (g->extrasrclen * (g->srcid != 1)) {
        if (g->srcid > 1) {
            src = (const mcxb_source*)(g->gproperty + g->maxmedia + 1 + (g->srcid - 2) * 4);
        } else {
            ppath[g->w0offset - 1] = (float)((int)(rand_uniform01(t) * JUST_BELOW_ONE * (float)(g->extrasrclen + 1)) + 1);

            if ((int)ppath[g->w0offset - 1] > 1) {
                src = (const mcxb_source*)(g->gproperty + g->maxmedia + 1 + ((int)(ppath[g->w0offset - 1] - 2.f)) * 4);
            }
        }
    }

    if (g->replay && g->srcid >= 1) {      /* :1614-1616 */
        (void)rand_uniform01(t);
    }

    /* :1619, as written: a bitwise AND of the source count with the truth value of srcid < 0 */
    const int cur_src_id = (g->extrasrclen & (uint32_t)(g->srcid < 0)) ? (int)ppath[g->w0offset - 1] : 0;
    ppath += g->partialdata;
    const f4 pos = {src->pos.x, src->pos.y, src->pos.z, src->pos.w};
    const f4 dir = {src->dir.x, src->dir.y, src->dir.z, src->dir.w};
    const f4 p1 = {src->param1.x, src->param1.y, src->param1.z, src->param1.w};
    const f4 p2 = {src->param2.x, src->param2.y, src->param2.z, src->param2.w};
    const int st = g->srctype;

    do {
        *p = pos;
        *v = dir;
        f->x = 0.f;
        f->y = 0.f;
        f->z = g->minaccumtime;
        it->idx1d = as_uint(p2.z);
        it->mediaid = as_uint(p2.w);

        if (g->mediaformat == 97) {
            it->svbits = 0u;           /* SV_CLEAR, :1640-1642 */
        }

        if (g->maxpolmedia > 0) {      /* :1633-1638 */
            it->si = g->s0.x;
            it->sq = g->s0.y;
            it->su = g->s0.z;
            it->sv = g->s0.w;
        }

        prop->x = pos.x;
        prop->y = pos.y;
        prop->z = pos.z;
        prop->w = 0.f;

        if (st == MCXB_SRC_PENCIL) {
            /* position, direction and launch voxel come straight from the source record (:1647-1653) */
        } else if (st == MCXB_SRC_PLANAR || st == MCXB_SRC_PATTERN || st == MCXB_SRC_PATTERN3D || st == MCXB_SRC_FOURIER || st == MCXB_SRC_PENCILARRAY) {
            /* :1656-1764 */
            const float rx = rand_uniform01(t);
            const float ry = rand_uniform01(t);
            float rz = 0.f;

            if (st == MCXB_SRC_PATTERN3D) {
                rz = rand_uniform01(t);
                p->x = p->x + rx * p1.x;
                p->y = p->y + ry * p1.y;
                p->z = p->z + rz * p1.z;
            }

            /* in the OpenCL build the `else` between the two offsets exists only under __NVCC__ (:1675-1678), so a
             * pattern3d packet receives the planar offset as well */
            p->x = p->x + rx * p1.x + ry * p2.x;
            p->y = p->y + rx * p1.y + ry * p2.y;
            p->z = p->z + rx * p1.z + ry * p2.z;

            if (st == MCXB_SRC_PATTERN || st == MCXB_SRC_PATTERN3D) {
                uint32_t cell;

                if (st == MCXB_SRC_PATTERN) {
                    cell = (uint32_t)((int)(ry * JUST_BELOW_ONE * p2.w) * (int)(p1.w) + (int)(rx * JUST_BELOW_ONE * p1.w));
                } else {
                    cell = (uint32_t)((int)(rz * JUST_BELOW_ONE * p1.z) * (int)(p1.y) * (int)(p1.x) +
                                      (int)(ry * JUST_BELOW_ONE * p1.y) * (int)(p1.x) + (int)(rx * JUST_BELOW_ONE * p1.x));
                }

                if (g->srcnum <= 1) {
                    p->w = pos.w * g->srcpattern[cell];
                    ppath[4] = p->w;
                } else {
                    memcpy(ppath + 2, &cell, 4);

                    for (uint32_t i = 0; i < g->srcnum; i++) {
                        ppath[i + 4] = g->srcpattern[(size_t)cell * g->srcnum + i];
                    }

                    p->w = 1.f;
                }
            } else if (st == MCXB_SRC_FOURIER) {
                p->w = pos.w * (cosf((floorf(p1.w) * rx + floorf(p2.w) * ry + p1.w - floorf(p1.w)) * TWO_PI) * (1.f - p2.w + floorf(p2.w)) + 1.f) * 0.5f;
            } else if (st == MCXB_SRC_PENCILARRAY) {
                p->x = pos.x + floorf(rx * p1.w) * p1.x / (p1.w - 1.f) + floorf(ry * p2.w) * p2.x / (p2.w - 1.f);
                p->y = pos.y + floorf(rx * p1.w) * p1.y / (p1.w - 1.f) + floorf(ry * p2.w) * p2.y / (p2.w - 1.f);
                p->z = pos.z + floorf(rx * p1.w) * p1.z / (p1.w - 1.f) + floorf(ry * p2.w) * p2.z / (p2.w - 1.f);
            }

            locate(g, it);
            prop->x = prop->x + (p1.x + p2.x) * 0.5f;
            prop->y = prop->y + (p1.y + p2.y) * 0.5f;
            prop->z = prop->z + (p1.z + p2.z) * 0.5f;
            prop->w = 0.f;
        } else if (st == MCXB_SRC_FOURIERX || st == MCXB_SRC_FOURIERX2D) {
            /* :1767-1814 */
            const float rx = rand_uniform01(t);
            const float ry = rand_uniform01(t);
            f4 v2 = p1;
            v2.w *= rsqrt_(p1.x * p1.x + p1.y * p1.y + p1.z * p1.z);
            v2.x = v2.w * (dir.y * p1.z - dir.z * p1.y);
            v2.y = v2.w * (dir.z * p1.x - dir.x * p1.z);
            v2.z = v2.w * (dir.x * p1.y - dir.y * p1.x);
            p->x = p->x + rx * p1.x + ry * v2.x;
            p->y = p->y + rx * p1.y + ry * v2.y;
            p->z = p->z + rx * p1.z + ry * v2.z;

            if (st == MCXB_SRC_FOURIERX2D) {
                p->w = pos.w * (sinf((p2.x * rx + p2.z) * TWO_PI) * sinf((p2.y * ry + p2.w) * TWO_PI) + 1.f) * 0.5f;
            } else {
                p->w = pos.w * (cosf((p2.x * rx + p2.y * ry + p2.z) * TWO_PI) * (1.f - p2.w) + 1.f) * 0.5f;
            }

            locate(g, it);
            prop->x = prop->x + (p1.x + v2.x) * 0.5f;
            prop->y = prop->y + (p1.y + v2.y) * 0.5f;
            prop->z = prop->z + (p1.z + v2.z) * 0.5f;
            prop->w = 0.f;
        } else if (st == MCXB_SRC_HYPERBOLOID_GAUSSIAN) {
            /* :1817-1853 */
            float sphi, cphi;
            float r = TWO_PI * rand_uniform01(t);
            sphi = sinf(r);
            cphi = cosf(r);
            r = sqrtf(0.5f * rand_next_scatlen(s, t)) * p1.x;
            prop->x = -p1.y / p1.z;
            prop->y = rsqrt_(r * r + p1.z * p1.z);
            p->x = r * (cphi - prop->x * sphi);
            p->y = r * (sphi + prop->x * cphi);
            p->z = 0.f;
            {
                const f4 d = {-r* sphi * prop->y, r* cphi * prop->y, p1.z * prop->y, 0.f};
                *prop = d;
            }

            if (v->z > -1.f + EPS && v->z < 1.f - EPS) {
                r = 1.f - v->z * v->z;
                const float stheta = sqrtf(r);
                r = rsqrt_(r);
                cphi = v->x * r;
                sphi = v->y * r;
                const f4 np = {p->x* cphi* v->z - p->y * sphi, p->x* sphi* v->z + p->y * cphi, -p->x * stheta, p->w};
                const f4 nv = {prop->x* cphi* v->z - prop->y* sphi + prop->z* cphi * stheta,
                               prop->x* sphi* v->z + prop->y* cphi + prop->z* sphi * stheta,
                               -prop->x* stheta + prop->z * v->z, v->w
                              };
                *p = np;
                *v = nv;
            } else {
                const f4 nv = {prop->x, prop->y, (v->z > 0.f) ? prop->z : -prop->z, v->w};
                *v = nv;
            }

            p->x = p->x + pos.x;
            p->y = p->y + pos.y;
            p->z = p->z + pos.z;
            prop->x = pos.x;
            prop->y = pos.y;
            prop->z = pos.z;
            prop->w = 0.f;
            it->Lmove = 0.f;
        } else if (st == MCXB_SRC_DISK || st == MCXB_SRC_GAUSSIAN || st == MCXB_SRC_RING) {
            /* :1856-1936 */
            float sphi, cphi, phi, r;

            if (st != MCXB_SRC_GAUSSIAN) {
                if (p1.z > 0.f || p1.w > 0.f) {
                    phi = fabsf(p1.z - p1.w) * rand_uniform01(t) + fminf(p1.z, p1.w);
                } else {
                    phi = TWO_PI * rand_uniform01(t);
                }
            } else {
                phi = TWO_PI * rand_uniform01(t);
            }

            sphi = sinf(phi);
            cphi = cosf(phi);

            if (st != MCXB_SRC_GAUSSIAN) {
                r = sqrtf(rand_uniform01(t) * fabsf(p1.x * p1.x - p1.y * p1.y) + p1.y * p1.y);
            } else if (fabsf(dir.w) < 1e-5f || fabsf(p1.y) < 1e-5f) {
                r = sqrtf(-0.5f * counted_log(s, rand_uniform01(t))) * p1.x;
            } else {
                r = p1.x * p1.x * ONE_PI / p1.y;
                r = sqrtf(-0.5f * counted_log(s, rand_uniform01(t)) * (1.f + (dir.w * dir.w / (r * r)))) * p1.x;
            }

            if (v->z > -1.f + EPS && v->z < 1.f - EPS) {
                const float tmp0 = 1.f - v->z * v->z;
                const float tmp1 = r * rsqrt_(tmp0);
                const float nx = p->x + tmp1 * (v->x * v->z * cphi - v->y * sphi);
                const float ny = p->y + tmp1 * (v->y * v->z * cphi + v->x * sphi);
                const float nz = p->z - tmp1 * tmp0 * cphi;
                p->x = nx;
                p->y = ny;
                p->z = nz;
            } else {
                p->x += r * cphi;
                p->y += r * sphi;
            }

            locate(g, it);
        } else if (st == MCXB_SRC_CONE || st == MCXB_SRC_ISOTROPIC || st == MCXB_SRC_ARCSINE) {
            /* :1939-1980 */
            float ang, stheta, ctheta, sphi, cphi;
            ang = TWO_PI * rand_uniform01(t);
            sphi = sinf(ang);
            cphi = cosf(ang);

            if (st == MCXB_SRC_CONE) {
                do {
                    ang = (p1.y > 0) ? TWO_PI * rand_uniform01(t) : acosf(2.f * rand_uniform01(t) - 1.f);
                } while (ang > p1.x);
            } else if (st == MCXB_SRC_ISOTROPIC) {
                ang = acosf(2.f * rand_uniform01(t) - 1.f);
            } else {
                ang = ONE_PI * rand_uniform01(t);
            }

            stheta = sinf(ang);
            ctheta = cosf(ang);
            rotatevector(v, stheta, ctheta, sphi, cphi);
            it->Lmove = 0.f;
        } else if (st == MCXB_SRC_ZGAUSSIAN) {
            /* :1983-1996 */
            float ang, stheta, ctheta, sphi, cphi;
            ang = TWO_PI * rand_uniform01(t);
            sphi = sinf(ang);
            cphi = cosf(ang);
            {
                const float a = sqrtf(-2.f * counted_log(s, rand_uniform01(t)));
                ang = a * (1.f - 2.f * rand_uniform01(t)) * p1.x;
            }
            stheta = sinf(ang);
            ctheta = cosf(ang);
            rotatevector(v, stheta, ctheta, sphi, cphi);
            it->Lmove = 0.f;
        } else if (st == MCXB_SRC_LINE || st == MCXB_SRC_SLIT) {
            /* :1999-2084 */
            float r_l = rand_uniform01(t);
            p->x = p->x + r_l * p1.x;
            p->y = p->y + r_l * p1.y;
            p->z = p->z + r_l * p1.z;

            if (st == MCXB_SRC_LINE) {
                float sphi_l, cphi_l;
                r_l = rsqrt_(p1.x * p1.x + p1.y * p1.y + p1.z * p1.z);

                if (p2.x > 0.f) {
                    const float ax = p1.x * r_l, ay = p1.y * r_l, az = p1.z * r_l;
                    const float vdotaxis = v->x * ax + v->y * ay + v->z * az;
                    v->x = v->x - vdotaxis * ax;
                    v->y = v->y - vdotaxis * ay;
                    v->z = v->z - vdotaxis * az;
                    const float vnorm = rsqrt_(v->x * v->x + v->y * v->y + v->z * v->z);
                    v->x = v->x * vnorm;
                    v->y = v->y * vnorm;
                    v->z = v->z * vnorm;
                    r_l = p2.x * (2.f * rand_uniform01(t) - 1.f);
                    sphi_l = sinf(r_l);
                    cphi_l = cosf(r_l);
                    rotate_perpendicular_vector(v, ax, ay, az, sphi_l, cphi_l);
                } else {
                    v->x = p1.x * r_l;
                    v->y = p1.y * r_l;
                    v->z = p1.z * r_l;
                    r_l = TWO_PI * rand_uniform01(t);
                    sphi_l = sinf(r_l);
                    cphi_l = cosf(r_l);
                    rotatevector(v, 1.f, 0.f, sphi_l, cphi_l);
                }
            } else if (p2.x > 0.f || p2.y > 0.f) {
                float sphi_s, cphi_s;
                r_l = TWO_PI * rand_uniform01(t);
                sphi_s = sinf(r_l);
                cphi_s = cosf(r_l);
                r_l = sqrtf(2.f * rand_next_scatlen(s, t));
                cphi_s *= p2.x * r_l;
                sphi_s *= p2.y * r_l;
                sphi_s *= rsqrt_(p1.x * p1.x + p1.y * p1.y + p1.z * p1.z);
                const f4 q = {p1.y* v->z - p1.z * v->y, p1.z* v->x - p1.x * v->z, p1.x* v->y - p1.y * v->x, 0.f};
                cphi_s *= rsqrt_(q.x * q.x + q.y * q.y + q.z * q.z);
                v->x += cphi_s * q.x + sphi_s * p1.x;
                v->y += cphi_s * q.y + sphi_s * p1.y;
                v->z += cphi_s * q.z + sphi_s * p1.z;
                r_l = rsqrt_(v->x * v->x + v->y * v->y + v->z * v->z);
                v->x *= r_l;
                v->y *= r_l;
                v->z *= r_l;
            }

            locate(g, it);
            prop->x = pos.x + p1.x * 0.5f;
            prop->y = pos.y + p1.y * 0.5f;
            prop->z = pos.z + p1.z * 0.5f;
            prop->w = 0.f;
            it->Lmove = -1.f;         /* both variants leave the focal-length rule enabled (:2082) */
        }

        if (fabsf(p->w) <= g->minenergy) {
            continue;                 /* :2094 */
        }

        /* launch angle (:2102-2153) */
        if (g->nangle) {
            float ang, stheta, ctheta, sphi, cphi;
            const float* at = g->sharedtab + g->nphaselen;

            if (dir.w > 0.f) {
                ang = fminf(rand_uniform01(t) * (float)g->nangle, (float)g->nangle - EPS);
                cphi = at[(int)ang];
            } else {
                ang = fminf(rand_uniform01(t) * (float)(g->nangle - 1), (float)(g->nangle - 1) - EPS);
                sphi = ang - (float)((int)ang);
                cphi = ((1.f - sphi) * at[(uint32_t)ang >= g->nangle - 1 ? g->nangle - 1 : (uint32_t)(int)ang] +
                        sphi * at[(uint32_t)ang + 1 >= g->nangle - 1 ? g->nangle - 1 : (uint32_t)((int)ang + 1)]);
            }

            cphi *= ONE_PI;
            stheta = sinf(cphi);
            ctheta = cosf(cphi);
            ang = TWO_PI * rand_uniform01(t);
            sphi = sinf(ang);
            cphi = cosf(ang);

            if (dir.w < 1.5f && dir.w >= 0.f) {
                *v = dir;
            }

            rotatevector(v, stheta, ctheta, sphi, cphi);
        } else if (it->Lmove < 0.f) {
            if (isnan(dir.w)) {
                float ang, stheta, ctheta, sphi, cphi;
                ang = TWO_PI * rand_uniform01(t);
                sphi = sinf(ang);
                cphi = cosf(ang);
                ang = acosf(2.f * rand_uniform01(t) - 1.f);
                stheta = sinf(ang);
                ctheta = cosf(ang);
                rotatevector(v, stheta, ctheta, sphi, cphi);
            } else if (dir.w < 0.f && isinf(dir.w)) {
                float ang, stheta, ctheta, sphi, cphi;
                ang = TWO_PI * rand_uniform01(t);
                sphi = sinf(ang);
                cphi = cosf(ang);
                stheta = sqrtf(rand_uniform01(t));
                ctheta = sqrtf(1.f - stheta * stheta);
                rotatevector(v, stheta, ctheta, sphi, cphi);
            } else if (dir.w != 0.f) {
                float Rn2 = (float)((dir.w > 0.f) - (dir.w < 0.f));
                prop->x += dir.w * v->x;
                prop->y += dir.w * v->y;
                prop->z += dir.w * v->z;
                v->x = Rn2 * (prop->x - p->x);
                v->y = Rn2 * (prop->y - p->y);
                v->z = Rn2 * (prop->z - p->z);
                Rn2 = rsqrt_(v->x * v->x + v->y * v->y + v->z * v->z);
                v->x *= Rn2;
                v->y *= Rn2;
                v->z *= Rn2;
            }
        }

        /* adjoint runs (and srcid == -2) launch the detectors appended to the source list as disks of their radius around
         * the detector position, perpendicular to the launch direction (:2155-2183) */
        if ((g->outputtype >= 11 || g->srcid == -2) && cur_src_id > (int)(g->extrasrclen + 1) - (int)g->detnum) {
            const float phi_adj = TWO_PI * rand_uniform01(t);
            const float sphi_adj = sinf(phi_adj), cphi_adj = cosf(phi_adj);
            const float r_adj = sqrtf(rand_uniform01(t)) * p1.x;

            if (v->z > -1.f + EPS && v->z < 1.f - EPS) {
                const float tmp0_adj = 1.f - v->z * v->z;
                const float tmp1_adj = r_adj * rsqrt_(tmp0_adj);
                p->x += tmp1_adj * (v->x * v->z * cphi_adj - v->y * sphi_adj);
                p->y += tmp1_adj * (v->y * v->z * cphi_adj + v->x * sphi_adj);
                p->z -= tmp1_adj * tmp0_adj * cphi_adj;
            } else {
                p->x += r_adj * cphi_adj;
                p->y += r_adj * sphi_adj;
            }

            it->idx1d = (uint32_t)((int)floorf(p->z) * (int)g->dimxy + (int)floorf(p->y) * (int)g->dimx + (int)floorf(p->x));

            if (p->x < 0.f || p->y < 0.f || p->z < 0.f || p->x >= g->maxidx.x || p->y >= g->maxidx.y || p->z >= g->maxidx.z) {
                it->mediaid = 0;
            } else {
                it->mediaid = g->media[it->idx1d];
            }
        }

        /* a packet launched in a zero voxel is marched to the surface (:2188-2195) */
        if ((it->mediaid & MED_MASK) == 0) {
            const int idx = skipvoid(g, s, p, v, f, &it->flipdir);

            if (idx >= 0) {
                it->idx1d = (uint32_t)idx;
                it->mediaid = g->media[it->idx1d];
            }
        }

        set_voxel_from_pos(&it->flipdir, p);
        it->w0 += 1.f;

        if (it->w0 > (float)g->maxvoidstep) {
            return -1;
        }
    } while ((it->mediaid & MED_MASK) == 0 || fabsf(p->w) <= g->minenergy);

    /* :2212-2255 */
    f->w += 1.f;
    *prop = g->gproperty[1];

    if (g->mediaformat == 97) {
        update_property_svmc(g, it, prop, it->mediaid, it->idx1d, p);
    } else {
        update_property(g, prop, it->mediaid);
    }

    if (g->debuglevel & (2u | 8u)) {           /* the start of a trajectory (:2243-2249) */
        savedebugdata(g, s, p, (uint32_t)f->w + (uint32_t)it->threadid * g->threadphoton + (uint32_t)(it->threadid < g->oddphoton ? it->threadid : 0), (int)ppath[3]);
    }

    ppath[1] += p->w;
    it->w0 = p->w;
    ppath[2] = (g->srcnum > 1) ? ppath[2] : p->w;
    v->w = EPS;
    it->Lmove = 0.f;
    return 0;
}

/* ------------------------------------------------------------------------ mcx_main_loop :2307-3307 */

static void work_item(const param_t* g, sink_t* s, int idx, uint32_t* n_seed, float* genergy, uint32_t global_size) {
    item_t it;
    memset(&it, 0, sizeof(it));
    it.si = 1.f;             /* Stokes s = {1, 0, 0, 0} (:2341) */
    float ppath_store[64];
    const uint32_t ppathlen = g->w0offset + g->srcnum;
    float* ppath = (ppathlen <= 64) ? ppath_store : (float*)malloc(sizeof(float) * ppathlen);
    it.p.w = NAN;
    it.mediaid = as_uint(g->cfg->src.param2.w);
    it.flipdir.w = -1;
    it.ppath = ppath;
    it.threadid = idx;
    uint32_t idx1dold, mediaidold = 0, isdet = 0;
    float pathlen = 0.f, n1;
    float w_re = 0.f, w_im = 0.f, w0_re = 0.f, w0_im = 0.f;      /* complex weight of an RF forward run (:2338) */
    const int rf = g->omega > 0.f;
    const int svmc = g->mediaformat == 97;      /* MEDIA_2LABEL_SPLIT */
    int testint = 0, hitintf = 0;               /* :2345-2346 */
    f4* p = &it.p;
    f4* v = &it.v;
    f4* f = &it.f;
    f4* prop = &it.prop;
    s4* flipdir = &it.flipdir;
    uint64_t* t = it.t;

    if ((uint32_t)idx >= g->threadphoton * global_size + (uint32_t)g->oddphoton) {
        goto done;       /* :2387 */
    }

    for (uint32_t i = 0; i < ppathlen; i++) {
        ppath[i] = 0.f;
    }

    ppath[g->partialdata] = genergy[idx << 1];
    ppath[g->partialdata + 1] = genergy[(idx << 1) + 1];
    rng_init(t, n_seed + (size_t)idx * 4);

    if (g->debuglevel & 1u) {
        /* -D R: the field is filled with the uniform draws of every stream (:2408-2414) */
        for (uint32_t i = (uint32_t)idx; i < g->fieldlen; i += global_size) {
            s->field[i] = rand_uniform01(t);
        }

        goto done;
    }

    if (launchnewphoton(g, s, &it, 0)) {
        n_seed[idx] = NO_LAUNCH;
        goto done;
    }

    isdet = it.mediaid & DET_MASK;
    it.mediaid &= MED_MASK;

    if (rf) {      /* :2427-2430 */
        w_re = w0_re = p->w;
        w_im = w0_im = 0.f;
    }

    while (f->w <= (float)(g->threadphoton + (idx < g->oddphoton))) {
        /* ---- scattering (:2446-2649) ---- */
        if (f->x <= 0.f) {
            f->x = rand_next_scatlen(s, t);

            if (v->w != EPS) {
                float cphi = 1.f, sphi = 0.f, theta, stheta, ctheta;
                float tmp0 = 0.f;

                if (g->maxpolmedia > 0 && !g->is2d) {
                    /* rejection sampling of (theta, phi) from the phase function of the CURRENT Stokes vector (:2454-2468) */
                    const uint32_t ipol = (uint32_t)NANGLES * ((it.mediaid & MED_MASK) - 1);
                    float I0, I, sin2phi, cos2phi;

                    do {
                        theta = acosf(2.f * rand_uniform01(t) - 1.f);
                        tmp0 = TWO_PI * rand_uniform01(t);
                        sin2phi = sinf(2.f * tmp0);
                        cos2phi = cosf(2.f * tmp0);
                        I0 = g->smatrix[ipol].x * it.si + g->smatrix[ipol].y * (it.sq * cos2phi + it.su * sin2phi);
                        const uint32_t ithedeg = (uint32_t)(theta * NANGLES * (R_PI - EPS));
                        I = g->smatrix[ipol + ithedeg].x * it.si + g->smatrix[ipol + ithedeg].y * (it.sq * cos2phi + it.su * sin2phi);
                    } while (rand_uniform01(t) * I0 >= I);

                    sphi = sinf(tmp0);
                    cphi = cosf(tmp0);
                    stheta = sinf(theta);
                    ctheta = cosf(theta);
                } else {
                if (!g->is2d) {
                    tmp0 = TWO_PI * rand_uniform01(t);
                    sphi = sinf(tmp0);
                    cphi = cosf(tmp0);
                }

                if (g->nphase > 2) {
                    tmp0 = rand_uniform01(t) * (float)(g->nphase - 1);
                    theta = tmp0 - (float)((int)tmp0);
                    tmp0 = (1.f - theta) * g->sharedtab[(uint32_t)tmp0 >= g->nphase ? g->nphase - 1 : (uint32_t)(int)tmp0] +
                           theta * g->sharedtab[(uint32_t)tmp0 + 1 >= g->nphase ? g->nphase - 1 : (uint32_t)((int)tmp0 + 1)];
                    theta = acosf(tmp0);
                    stheta = sinf(theta);
                    ctheta = tmp0;
                } else {
                    tmp0 = (v->w > (float)g->gscatter) ? 0.f : prop->z;

                    if (fabsf(tmp0) > EPS) {
                        /* Henyey-Greenstein inverse CDF (:2487-2490) */
                        tmp0 = (1.f - prop->z * prop->z) / (1.f - prop->z + 2.f * prop->z * rand_uniform01(t));
                        tmp0 *= tmp0;
                        tmp0 = (1.f + prop->z * prop->z - tmp0) / (2.f * prop->z);
                        tmp0 = fmaxf(-1.f, fminf(1.f, tmp0));
                        theta = acosf(tmp0);
                        stheta = sinf(theta);
                        ctheta = tmp0;
                    } else {
                        theta = acosf(2.f * rand_uniform01(t) - 1.f);
                        stheta = sinf(theta);
                        ctheta = cosf(theta);
                    }
                }
                }

                if (g->savedet) {
                    /* split-voxel media count by the part of the voxel the packet is in, and not at all in an empty part (:2504-2545) */
                    const uint32_t lab = svmc ? SV_CURLABEL(it.svbits) : (it.mediaid & MED_MASK);

                    if ((g->savedetflag & 0x02u) && (!svmc || lab > 0)) {
                        uint32_t c;
                        memcpy(&c, ppath + lab - 1, 4);
                        c++;
                        memcpy(ppath + lab - 1, &c, 4);
                    }

                    if ((g->savedetflag & 0x08u) && (!svmc || lab > 0)) {
                        ppath[g->maxmedia * ((g->savedetflag >> 1 & 1u) + (g->savedetflag >> 2 & 1u)) + lab - 1] += 1.f - ctheta;
                    }
                }

                const f4 olddir = *v;

                if (g->is2d) {
                    rotatevector2d(v, (rand_uniform01(t) > 0.5f ? stheta : -stheta), ctheta, (int)g->is2d);
                } else {
                    rotatevector(v, stheta, ctheta, sphi, cphi);
                }

                v->w += 1.f;

                if (g->maxpolmedia > 0) {      /* :2562-2565 */
                    updatestokes(g, &it, theta, tmp0, &olddir, v);
                }

                /* replay: weighted scattering counts / momentum transfer at the scattering site (:2568-2612), with the record
                 * index as the reference writes it (f.w, not f.w - 1) and its way of moving an overflowing sum to the shadow half */
                if (g->outputtype == otWP || g->outputtype == otDCS || g->outputtype == otWPTOF) {
                    const int rec = idx * (int)g->threadphoton + (idx < g->oddphoton - 1 ? idx : g->oddphoton - 1) + (int)f->w;
                    int tshift = rec;
                    tmp0 = (g->outputtype == otDCS) ? (1.f - ctheta) : 1.f;

                    if (g->outputtype == otWPTOF) {
                        tmp0 = g->rtof[rec];
                        tmp0 *= g->rweight[rec];
                    } else {
                        tmp0 *= g->rweight[rec];
                    }

                    tshift = (int)floorf((g->rtof[tshift] - g->twin0) * g->Rtstep) +
                             ((g->replaydet == -1) ? (((g->rdetid[tshift] & 0xFFFF) - 1) * (int)g->maxgate) : 0);

                    if (g->extrasrclen * (g->srcid < 0)) {
                        tshift += ((int)ppath[g->w0offset - 1] - 1) * (int)((g->replaydet == -1) ? g->detnum : 1u) * (int)g->maxgate;
                    }

                    tshift = ((int)g->maxgate - 1 < tshift) ? (int)g->maxgate - 1 : tshift;
                    float* at = s->field + it.idx1d + (size_t)tshift * g->dimxyz;
                    const float oldval = atomicadd(s, at, tmp0);

                    if (fabsf(oldval) > MAX_ACCUM) {
                        if (atomicadd(s, at, -oldval) < 0.f) {
                            atomicadd(s, at, oldval);
                        } else {
                            atomicadd(s, at + g->fieldlen, oldval);
                        }
                    }
                }

                if (g->debuglevel & (2u | 8u)) {   /* every scattering site (:2625-2632) */
                    savedebugdata(g, s, p, (uint32_t)f->w + (uint32_t)idx * g->threadphoton + (uint32_t)(idx < g->oddphoton ? idx : 0), (int)ppath[g->w0offset - 1]);
                }
            }

            v->w = (float)(int)v->w;

            if (svmc) {
                testint = 1;           /* :2638-2648 */
            }
        }

        /* ---- one ray segment (:2652-2765) ---- */
        n1 = prop->w;

        if (svmc) {
            update_property_svmc(g, &it, prop, it.mediaid, it.idx1d, p);
        } else {
            update_property(g, prop, it.mediaid);
        }

        f->z = hitgrid(s, p, v, flipdir);
        float slen = f->z * prop->y * (v->w + 1.f > (float)g->gscatter ? (1.f - prop->z) : 1.f);
        slen = fminf(slen, f->x);
        f->z = slen / (prop->y * (v->w + 1.f > (float)g->gscatter ? (1.f - prop->z) : 1.f));

        if (svmc && SV_ISSPLIT(it.svbits) && testint) {       /* :2684-2699 */
            float tmplen = f->z;
            hitintf = ray_plane_intersect(g, &it, prop, &tmplen, &slen);
            f->z = tmplen;
        } else {
            hitintf = 0;
        }

        pathlen += f->z;
        p->x = p->x + f->z * v->x;
        p->y = p->y + f->z * v->y;
        p->z = p->z + f->z * v->z;

        if (flipdir->w == 0) {
            flipdir->x += (slen == f->x || hitintf) ? 0 : (v->x > 0.f ? 1 : -1);
        }

        if (flipdir->w == 1) {
            flipdir->y += (slen == f->x || hitintf) ? 0 : (v->y > 0.f ? 1 : -1);
        }

        if (flipdir->w == 2) {
            flipdir->z += (slen == f->x || hitintf) ? 0 : (v->z > 0.f ? 1 : -1);
        }

        if (rf) {
            /* w <- w exp[-(mua + i omega n / c0) ds]; the magnitude drives roulette and the ledger (:2750-2760) */
            const float rf_atten = expf(-prop->x * f->z);
            const float ang = g->omega * prop->w * g->oneoverc0 * f->z;
            const float rf_sin = sinf(ang), rf_cos = cosf(ang);
            const float tmp_re = rf_atten * (w_re * rf_cos + w_im * rf_sin);
            const float tmp_im = rf_atten * (-w_re * rf_sin + w_im * rf_cos);
            w_re = tmp_re;
            w_im = tmp_im;
            p->w = sqrtf(w_re * w_re + w_im * w_im);
        } else {
            p->w *= expf(-prop->x * f->z);
        }
        f->x -= slen;
        f->y += f->z * prop->w * g->oneoverc0;

        if (g->savedet && (g->savedetflag & 0x04u)) {
            if (!svmc) {
                ppath[g->maxmedia * (g->savedetflag >> 1 & 1u) + (it.mediaid & MED_MASK) - 1] += f->z;
            } else if (SV_CURLABEL(it.svbits) > 0) {      /* :2772-2782 */
                ppath[g->maxmedia * (g->savedetflag >> 1 & 1u) + SV_CURLABEL(it.svbits) - 1] += f->z;
            }
        }

        /* ---- new voxel (:2796-2811) ---- */
        mediaidold = it.mediaid | isdet;
        idx1dold = it.idx1d;
        it.idx1d = voxel_index(g, flipdir);

        if (!in_grid(g, flipdir)) {
            it.mediaid = 0;
            it.idx1d = (flipdir->x < 0 || flipdir->y < 0 || flipdir->z < 0) ? OUTSIDE_VOLUME_MIN : OUTSIDE_VOLUME_MAX;
            isdet = g->bc[(it.idx1d == OUTSIDE_VOLUME_MAX) * 3 + flipdir->w];
            isdet = ((isdet & 0xF) == bcUnknown) ? (g->doreflect ? bcReflect : bcAbsorb) : isdet;
        } else {
            it.mediaid = g->media[it.idx1d];
            isdet = it.mediaid & DET_MASK;
            it.mediaid &= MED_MASK;
        }

        /* ---- deposit on leaving a voxel (:2816-2929) ---- */
        if ((it.idx1d != idx1dold || hitintf) && idx1dold < g->dimxyz && mediaidold) {
            if (g->save2pt && f->y >= g->twin0 && f->y < g->twin1) {
                float weight = 0.f, weight_im = 0.f;
                int tshift = (int)floorf((f->y - g->twin0) * g->Rtstep);

                if (rf) {
                    /* the complex quotient (w0 - w) / (mua + i omega n / c0) (:2833-2841) */
                    const float dw_re = w0_re - w_re, dw_im = w0_im - w_im;
                    const float a_im = g->omega * prop->w * g->oneoverc0;
                    const float a_mag2 = prop->x * prop->x + a_im * a_im;
                    weight = (a_mag2 < EPS) ? (w0_re * pathlen) : (dw_re * prop->x + dw_im * a_im) / a_mag2;
                    weight_im = (a_mag2 < EPS) ? (w0_im * pathlen) : (dw_im * prop->x - dw_re * a_im) / a_mag2;
                } else if (g->outputtype == otEnergy) {
                    weight = it.w0 - p->w;
                } else if (g->outputtype == otFluence || g->outputtype == otFlux || g->outputtype >= 11) {      /* 11..16: the adjoint types (:2844) */
                    weight = (prop->x < EPS) ? (it.w0 * pathlen) : ((it.w0 - p->w) / (prop->x));
                } else if (g->replay) {
                    if (g->outputtype == otJacobian || g->outputtype == otWLTOF) {
                        /* w_i L binned by the DETECTED time of flight of record i (:2845-2858) */
                        const int rec = idx * (int)g->threadphoton + (idx < g->oddphoton - 1 ? idx : g->oddphoton - 1) + (int)f->w - 1;
                        weight = g->rweight[rec] * pathlen;
                        tshift = (int)floorf((g->rtof[rec] - g->twin0) * g->Rtstep) +
                                 ((g->replaydet == -1) ? (((g->rdetid[rec] & 0xFFFF) - 1) * (int)g->maxgate) : 0);

                        if (g->outputtype == otWLTOF) {
                            weight = weight * g->rtof[rec];
                        }
                    }
                } else if (g->outputtype == otL) {
                    weight = it.w0 * pathlen;
                }

                if (g->extrasrclen * (g->srcid < 0)) {
                    tshift += ((int)ppath[g->w0offset - 1] - 1) * (int)((g->replaydet == -1) ? g->detnum : 1u) * (int)g->maxgate;
                }

                if (fabsf(weight) > 0.f) {
                    if (g->srctype != MCXB_SRC_PATTERN && g->srctype != MCXB_SRC_PATTERN3D) {
                        const size_t at = (size_t)(idx1dold + (uint32_t)tshift * g->dimxyz);

                        if (!rf) {
                            deposit(g, s, at, weight);
                        } else {
                            /* :2882-2893 as written: the imaginary part goes to the third quarter of the buffer in the ELSE
                             * branch of the real part's spill, so the step that spills the real part loses it */
                            const float oldval = atomicadd(s, s->field + at, weight);

                            if (fabsf(oldval) > MAX_ACCUM) {
                                atomicadd(s, s->field + at, (oldval > 0.f) ? -MAX_ACCUM : MAX_ACCUM);
                                atomicadd(s, s->field + at + g->fieldlen, (oldval > 0.f) ? MAX_ACCUM : -MAX_ACCUM);
                            } else {
                                const float oldim = atomicadd(s, s->field + at + 2 * (size_t)g->fieldlen, weight_im);

                                if (fabsf(oldim) > MAX_ACCUM) {
                                    atomicadd(s, s->field + at + 2 * (size_t)g->fieldlen, (oldim > 0.f) ? -MAX_ACCUM : MAX_ACCUM);
                                    atomicadd(s, s->field + at + 3 * (size_t)g->fieldlen, (oldim > 0.f) ? MAX_ACCUM : -MAX_ACCUM);
                                }
                            }
                        }
                    } else {
                        for (uint32_t i = 0; i < g->srcnum; i++) {
                            if (fabsf(ppath[g->w0offset + i]) > 0.f) {
                                deposit(g, s, (size_t)(idx1dold + (uint32_t)tshift * g->dimxyz) * g->srcnum + i,
                                        (g->srcnum == 1) ? weight : weight * ppath[g->w0offset + i]);
                            }
                        }
                    }
                }
            }

            it.w0 = p->w;

            if (rf) {
                w0_re = w_re;
                w0_im = w_im;
            }

            pathlen = 0.f;
        } else {
            it.mediaid = mediaidold;     /* note: carries the detector / boundary bits of isdet along (:2928) */
        }

        if (svmc) {
            /* the tissue on the far side of a voxel face or of the interface (:2931-2949) */
            if (it.idx1d != idx1dold) {
                update_property_svmc(g, &it, prop, it.mediaid, it.idx1d, p);
                testint = 1;
            } else if (hitintf) {
                it.svnx = -it.svnx;
                it.svny = -it.svny;
                it.svnz = -it.svnz;
                it.svpd = -it.svpd;
                it.svbits ^= 0x20000u;
                testint = 0;
            }
        }

        /* ---- leave the domain, time out, or wrap around (:2957-3028) ---- */
        if ((it.mediaid == 0 && ((isdet & 0xF) == bcAbsorb || (isdet & 0xF) == bcCyclic || ((isdet & 0xF) == bcReflect && n1 == g->gproperty[0].w)))
                || (svmc && ((it.idx1d != idx1dold || hitintf) && !SV_ISUPPER(it.svbits) && !SV_LOWER(it.svbits)
                             && (!g->doreflect || n1 == g->gproperty[0].w)))
                || f->y > g->twin1) {
            if (isdet == bcCyclic) {
                if (flipdir->w == 0) {
                    p->x = mcx_nextafterf(rintf((it.idx1d == OUTSIDE_VOLUME_MIN) ? g->maxidx.x : 0.f), (v->x > 0.f) - (v->x < 0.f));
                    flipdir->x = (short)floorf(p->x);
                }

                if (flipdir->w == 1) {
                    p->y = mcx_nextafterf(rintf((it.idx1d == OUTSIDE_VOLUME_MIN) ? g->maxidx.y : 0.f), (v->y > 0.f) - (v->y < 0.f));
                    flipdir->y = (short)floorf(p->y);
                }

                if (flipdir->w == 2) {
                    p->z = mcx_nextafterf(rintf((it.idx1d == OUTSIDE_VOLUME_MIN) ? g->maxidx.z : 0.f), (v->z > 0.f) - (v->z < 0.f));
                    flipdir->z = (short)floorf(p->z);
                }

                if (in_grid(g, flipdir)) {
                    it.idx1d = voxel_index(g, flipdir);
                    it.mediaid = g->media[it.idx1d];
                    isdet = it.mediaid & DET_MASK;
                    it.mediaid &= MED_MASK;
                    continue;
                }
            }

            if (launchnewphoton(g, s, &it, (((it.idx1d == OUTSIDE_VOLUME_MAX && g->bc[9 + flipdir->w]) || (it.idx1d == OUTSIDE_VOLUME_MIN && g->bc[6 + flipdir->w]))
                                            ? OUTSIDE_VOLUME_MIN : (mediaidold & DET_MASK)))) {
                break;
            }

            isdet = it.mediaid & DET_MASK;
            it.mediaid &= MED_MASK;

            if (rf) {      /* :3011-3014 */
                w_re = w0_re = p->w;
                w_im = w0_im = 0.f;
            }

            if (svmc) {
                testint = 1;           /* :3016-3026 */
            }

            continue;
        }

        /* ---- Russian roulette (:3031-3061) ---- */
        if (fabsf(p->w) < g->minenergy) {
            if (rand_uniform01(t) * ROULETTE_SIZE <= 1.f) {
                p->w *= ROULETTE_SIZE;

                if (rf) {
                    w_re *= ROULETTE_SIZE;
                    w_im *= ROULETTE_SIZE;
                    w0_re *= ROULETTE_SIZE;
                    w0_im *= ROULETTE_SIZE;
                }
            } else {
                if (launchnewphoton(g, s, &it, mediaidold & DET_MASK)) {
                    break;
                }

                isdet = it.mediaid & DET_MASK;
                it.mediaid &= MED_MASK;

                if (rf) {
                    w_re = w0_re = p->w;
                    w_im = w0_im = 0.f;
                }

                continue;
            }
        }

        /* ---- refractive-index mismatch (:3063-3297), compiled in under MCX_DO_REFLECTION ---- */
        if (g->doreflection) {
            if (svmc && hitintf) {
                /* Fresnel reflection / refraction at the interface inside a split voxel (:3074-3111) */
                if (g->gproperty[SV_LOWER(it.svbits)].w != g->gproperty[SV_UPPER(it.svbits)].w) {
                    it.svnx = -it.svnx;
                    it.svny = -it.svny;
                    it.svnz = -it.svnz;
                    it.svpd = -it.svpd;
                    float c0[3] = { v->x, v->y, v->z };

                    if (reflectray_svmc(g, &it, n1, c0, prop)) {
                        if (launchnewphoton(g, s, &it, mediaidold & DET_MASK)) {
                            break;
                        }

                        isdet = it.mediaid & DET_MASK;
                        it.mediaid &= MED_MASK;

                        if (rf) {
                            w_re = w0_re = p->w;
                            w_im = w0_im = 0.f;
                        }

                        testint = 1;
                        continue;
                    }

                    v->x = c0[0];
                    v->y = c0[1];
                    v->z = c0[2];
                } else {
                    *prop = g->gproperty[SV_CURLABEL(it.svbits)];
                }

            } else {
            if (!svmc) {
                update_property(g, prop, it.mediaid);
            }

            /* the index on the far side: the decoded one up to format 99, row 1 (row 0 outside) of the table from 100 on (:3147) */
            const float nfar = (g->mediaformat < 100) ? prop->w : g->gproperty[it.mediaid > 0 ? 1 : 0].w;

            if (((it.mediaid && g->doreflect)
                    || (it.mediaid == 0 && (((isdet & 0xF) == bcUnknown && g->doreflect) || ((isdet & 0xF) == bcReflect || (isdet & 0xF) == bcMirror))))
                    && (((isdet & 0xF) == bcMirror) || n1 != nfar)) {
                float Rtotal = 1.f;
                float cphi, sphi, stheta, ctheta;
                const float tmp0 = n1 * n1;
                const float tmp1 = prop->w * prop->w;
                cphi = fabsf((flipdir->w == 0) ? v->x : (flipdir->w == 1 ? v->y : v->z));
                sphi = 1.f - cphi * cphi;
                f->z = 1.f - tmp0 / tmp1 * sphi;

                if (f->z > 0.f && (isdet & 0xF) != bcMirror) {
                    ctheta = tmp0 * cphi * cphi + tmp1 * f->z;
                    stheta = 2.f * n1 * prop->w * cphi * sqrtf(f->z);
                    Rtotal = (ctheta - stheta) / (ctheta + stheta);
                    ctheta = tmp1 * cphi * cphi + tmp0 * f->z;
                    Rtotal = (Rtotal + (ctheta - stheta) / (ctheta + stheta)) * 0.5f;
                }

                if (Rtotal < 1.f && (!(it.mediaid == 0 && ((isdet & 0xF) == bcMirror))) && rand_uniform01(t) > Rtotal) {
                    transmit(v, n1, prop->w, flipdir->w);

                    if (it.mediaid == 0) {
                        if (launchnewphoton(g, s, &it, (((it.idx1d == OUTSIDE_VOLUME_MAX && g->bc[9 + flipdir->w]) || (it.idx1d == OUTSIDE_VOLUME_MIN && g->bc[6 + flipdir->w]))
                                                        ? OUTSIDE_VOLUME_MIN : (mediaidold & DET_MASK)))) {
                            break;
                        }

                        isdet = it.mediaid & DET_MASK;
                        it.mediaid &= MED_MASK;

                        if (rf) {      /* :3191-3194 */
                            w_re = w0_re = p->w;
                            w_im = w0_im = 0.f;
                        }

                        continue;
                    }
                } else {
                    /* mirror the direction and put the packet back on the face it came through (:3204-3213); the
                     * direction argument (v>0)-0.5f truncates to 0 when converted to int */
                    if (flipdir->w == 0) {
                        v->x = -v->x;
                        p->x = mcx_nextafterf(rintf(p->x), (int)((float)(v->x > 0.f) - 0.5f));
                        flipdir->x = (short)rintf(p->x);
                    } else if (flipdir->w == 1) {
                        v->y = -v->y;
                        p->y = mcx_nextafterf(rintf(p->y), (int)((float)(v->y > 0.f) - 0.5f));
                        flipdir->y = (short)rintf(p->y);
                    } else {
                        v->z = -v->z;
                        p->z = mcx_nextafterf(rintf(p->z), (int)((float)(v->z > 0.f) - 0.5f));
                        flipdir->z = (short)rintf(p->z);
                    }

                    it.idx1d = idx1dold;
                    /* the reference reads media[idx1dold] unguarded (:3212-3213); with split-voxel media idx1dold can be an
                     * OUTSIDE_VOLUME marker in rare walks (undefined there), where this restatement reads an empty word */
                    it.mediaid = (it.idx1d < g->dimxyz) ? (g->media[it.idx1d] & MED_MASK) : 0u;

                    if (svmc) {
                        /* back in the voxel it came from: which part of it? (:3215-3244) */
                        update_property_svmc(g, &it, prop, it.mediaid, it.idx1d, p);

                        if (SV_CURLABEL(it.svbits) == 0) {
                            if (launchnewphoton(g, s, &it, mediaidold & DET_MASK)) {
                                break;
                            }

                            isdet = it.mediaid & DET_MASK;
                            it.mediaid &= MED_MASK;

                            if (rf) {
                                w_re = w0_re = p->w;
                                w_im = w0_im = 0.f;
                            }

                            continue;
                        }
                    } else {
                        update_property(g, prop, it.mediaid);
                    }

                    n1 = prop->w;
                }
            }
            }

            if ((g->debuglevel & 8u) && (it.mediaid == 0 || it.idx1d == OUTSIDE_VOLUME_MIN || it.idx1d == OUTSIDE_VOLUME_MAX)) {
                goto done;       /* "should never happen" (:3287-3291): the work-item returns without its energy write-back */
            }
        }
    }

    /* ---- energy ledger (:3301-3302) ---- */
    genergy[idx << 1] = ppath[g->partialdata];
    genergy[(idx << 1) + 1] = ppath[g->partialdata + 1];
done:

    if (ppath != ppath_store) {
        free(ppath);
    }
}

/* ---------------------------------------------------------------------------------- host side */

/* src/mcx_host.cpp:696-700, 759-768: ONE glibc rand() stream, four 31-bit words per work-item */
void mcxo_seeds(int seed, uint64_t skip_records, uint64_t nrecords, uint32_t* out4) {
    srand(seed > 0 ? (unsigned)seed : 1u);

    for (uint64_t i = 0; i < skip_records * 4; i++) {
        (void)rand();
    }

    for (uint64_t i = 0; i < nrecords * 4; i++) {
        out4[i] = (uint32_t)rand();
    }
}

/* the reference's rule for compiling the reflection code in (src/mcx_host.cpp:945-956) */
static int needs_reflection(const mcxb_config* cfg) {
    int allabsorb = 1, allunknown = 1;

    for (int i = 0; i < 6; i++) {
        if (cfg->bc[i] != MCXB_BC_ABSORB) {
            allabsorb = 0;
        }

        if (cfg->bc[i] != 0) {
            allunknown = 0;
        }
    }

    if (cfg->bc[0] == 0) {     /* strcmp() stops at the first NUL */
        allunknown = 1;
        allabsorb = 0;
    }

    return cfg->isreflect || (!allabsorb && !allunknown);
}

int mcxo_run(const mcxb_config* cfg, uint32_t nthread, int hostthreads, mcxo_result* res) {
    if (cfg->abi_version != MCXB_ABI_VERSION || cfg->srctype < 0 || cfg->srctype >= 18 || nthread == 0) {
        return -1;
    }

    const int replay = cfg->replay_seed != NULL;
    const int sens = cfg->outputtype == otJacobian || cfg->outputtype == otWP || cfg->outputtype == otDCS || cfg->outputtype == otWLTOF ||
                     cfg->outputtype == otWPTOF;

    if ((sens && (!replay || !cfg->replay_weight || !cfg->replay_tof || (cfg->replaydet == -1 && !cfg->replay_detid))) ||
            (replay && (cfg->omega > 0.f || cfg->srcnum > 1))) {
        return -3;      /* the sensitivity outputs belong to a replay; RF replay is checked by oracle/_ref only (its phase factors are lost there) */
    }

    if (replay) {
        /* the reference maps work-item t to record t*threadphoton + min(t, oddphoton-1) + k (:1591): the identity only when a
         * work-item owns at most one photon, so a replay runs with nphoton + 1 work-items (like oracle/ref_driver.cpp) */
        nthread = (uint32_t)cfg->nphoton + 1;
    }

    const int continuous = cfg->mediaformat >= 99 && cfg->mediaformat <= 104;
    const int splitvox = cfg->mediaformat == 97;       /* MEDIA_2LABEL_SPLIT */

    if (splitvox && (cfg->isspecular > 0 || cfg->srcnum > 1 || cfg->polmedianum || cfg->replay_seed)) {
        return -3;      /* the reference indexes its media table with the whole media word under isspecular (:1425); the rest is untested there */
    }

    if ((cfg->mediaformat > 4 && !continuous && !splitvox) || (cfg->polmedianum && (!cfg->smatrix || continuous || cfg->medianum != cfg->polmedianum + 1)) || cfg->outputtype > 16 || cfg->outputtype == 6 || cfg->outputtype == 8) {
        return -3;      /* split-voxel / two-word media, RF replay and adjoint runs: checked by oracle/_ref only */
    }

    if (cfg->srcid < -2 || cfg->issaveref > 1) {
        return -3;      /* issaveref > 1 is a data race in the reference (:852-865): nothing to restate */
    }

    const int rfforward = cfg->omega > 0.f;

    if (rfforward && (cfg->srctype == MCXB_SRC_PATTERN || cfg->srctype == MCXB_SRC_PATTERN3D || (cfg->debuglevel & 1u))) {
        return -3;      /* the pattern builds of the reference have no imaginary deposit (:2902-2913) */
    }

    if (continuous && ((cfg->issavedet && (cfg->savedetflag & 0x0Eu)) || cfg->isspecular > 0 || cfg->srcnum > 1 ||
                       ((cfg->mediaformat == 103 || cfg->mediaformat == 104) && cfg->medianum < 3))) {
        return -3;      /* the reference indexes per-medium rows with the media WORD there (:1425, 2515, 2787): nothing defined to restate */
    }

    param_t g;
    memset(&g, 0, sizeof(g));
    g.cfg = cfg;
    g.dimx = cfg->dimx;
    g.dimxy = cfg->dimx * cfg->dimy;
    g.dimxyz = g.dimxy * cfg->dimz;
    g.maxgate = (uint32_t)((cfg->tend - cfg->tstart) / cfg->tstep + 0.5);       /* src/mcx_host.cpp:647 */
    g.mediaformat = (continuous || splitvox) ? cfg->mediaformat : 1u;
    g.omega = cfg->omega;
    g.replay = replay;
    g.replaydet = cfg->replaydet;
    g.rseed = (const uint64_t*)cfg->replay_seed;
    g.rweight = cfg->replay_weight;
    g.rtof = cfg->replay_tof;
    g.rdetid = cfg->replay_detid;
    g.maxpolmedia = cfg->smatrix ? cfg->polmedianum : 0;     /* src/mcx_host.cpp:522-523 */
    g.smatrix = (const f4*)cfg->smatrix;
    g.s0.x = cfg->srciquv.x;
    g.s0.y = cfg->srciquv.y;
    g.s0.z = cfg->srciquv.z;
    g.s0.w = cfg->srciquv.w;
    g.srcnum = cfg->srcnum ? cfg->srcnum : 1;
    const uint32_t nsrcvol = (cfg->srctype == MCXB_SRC_PATTERN || cfg->srctype == MCXB_SRC_PATTERN3D) ? g.srcnum
                             : ((cfg->srcid < 0) ? (cfg->extrasrclen + 1) : 1);
    /* src/mcx_host.cpp:684-689: one volume per detector when every detector is replayed at once */
    const uint32_t nrepvol = (replay && cfg->replaydet == -1) ? (cfg->detnum ? cfg->detnum : 1) : 1;
    const size_t fieldlen = (size_t)g.dimxyz * g.maxgate * nsrcvol * nrepvol;
    g.fieldlen = (uint32_t)fieldlen;
    g.maxidx.x = (float)cfg->dimx;
    g.maxidx.y = (float)cfg->dimy;
    g.maxidx.z = (float)cfg->dimz;
    g.twin0 = cfg->tstart;
    g.twin1 = cfg->tstart + cfg->tstep * (float)g.maxgate;
    g.oneoverc0 = R_C0 * cfg->unitinmm;
    g.Rtstep = 1.f / cfg->tstep;
    g.minenergy = cfg->minenergy;
    g.minaccumtime = cfg->unitinmm * R_C0 * cfg->unitinmm;                     /* src/mcx_host.cpp:515 */
    g.save2pt = (uint32_t)cfg->issave2pt;
    g.doreflect = (uint32_t)cfg->isreflect;
    g.savedet = (uint32_t)cfg->issavedet;
    g.maxdetphoton = cfg->maxdetphoton;
    g.maxmedia = cfg->medianum - 1;
    g.detnum = cfg->detnum;
    g.voidtime = cfg->voidtime;
    g.srctype = cfg->srctype;
    g.srcid = cfg->srcid;
    g.extrasrclen = cfg->extrasrclen;
    g.maxvoidstep = (uint32_t)cfg->maxvoidstep;
    g.issaveseed = cfg->issaveseed > 0;
    g.issaveref = (uint32_t)cfg->issaveref;
    g.isspecular = cfg->isspecular > 0;
    g.outputtype = (uint32_t)cfg->outputtype;
    g.threadphoton = (uint32_t)(cfg->nphoton / nthread);                       /* src/mcx_host.cpp:1011-1012 */
    g.oddphoton = (int)(cfg->nphoton - (uint64_t)g.threadphoton * nthread);
    g.debuglevel = cfg->debuglevel & (1u | 2u | 8u);       /* -D R, -D M, -D T */
    g.maxjumpdebug = cfg->maxjumpdebug;
    g.savedetflag = cfg->issavedet ? cfg->savedetflag : 0;
    g.partialdata = (cfg->medianum - 1) * ((g.savedetflag >> 1 & 1u) + (g.savedetflag >> 2 & 1u) + (g.savedetflag >> 3 & 1u));
    g.w0offset = g.partialdata + 4;
    g.reclen = g.partialdata + (g.savedetflag & 1u) + 3 * ((g.savedetflag >> 4 & 1u) + (g.savedetflag >> 5 & 1u)) + (g.savedetflag >> 6 & 1u) + 4 * (g.savedetflag >> 7 & 1u);
    g.gscatter = cfg->gscatter;
    g.is2d = (cfg->dimx == 1 ? 1 : (cfg->dimy == 1 ? 2 : (cfg->dimz == 1 ? 3 : 0)));

    if (g.is2d) {
        g.is2d = g.is2d * (((cfg->dimx > 1) + (cfg->dimy > 1) + (cfg->dimz > 1)) == 2);
    }

    g.nphase = cfg->invcdf ? cfg->nphase : 0;
    g.nphaselen = g.nphase + (g.nphase & 1);
    g.nangle = cfg->angleinvcdf ? cfg->nangle : 0;
    g.nanglelen = g.nangle + (g.nangle & 1);
    g.doreflection = needs_reflection(cfg);
    memcpy(g.bc, cfg->bc, 12);
    g.media = cfg->vol;
    g.srcpattern = cfg->srcpattern;

    f4* gproperty = (f4*)calloc(cfg->medianum + 4 * (size_t)cfg->extrasrclen + 1, sizeof(f4));
    memcpy(gproperty, cfg->prop, sizeof(f4) * cfg->medianum);

    if (cfg->extrasrclen) {
        memcpy(gproperty + cfg->medianum, cfg->srcdata, sizeof(mcxb_source) * cfg->extrasrclen);
    }

    g.gproperty = gproperty;
    g.gdetpos = (const f4*)cfg->detpos;
    float* sharedtab = (float*)calloc(g.nphaselen + g.nanglelen + 1, sizeof(float));

    if (g.nphase) {
        memcpy(sharedtab, cfg->invcdf, sizeof(float) * g.nphase);
    }

    if (g.nangle) {
        memcpy(sharedtab + g.nphaselen, cfg->angleinvcdf, sizeof(float) * g.nangle);
    }

    g.sharedtab = sharedtab;

    uint32_t* seeds = (uint32_t*)calloc(4 * (size_t)nthread + 4, sizeof(uint32_t));

    if (replay) {
        /* gseed holds the recorded RNG states (src/mcx_host.cpp:724); gpu_rng_init still reads one record per work-item */
        memcpy(seeds, cfg->replay_seed, 16 * (size_t)cfg->nphoton);
    } else {
        mcxo_seeds(cfg->seed, cfg->seed_skip, nthread, seeds);
    }
    float* genergy = (float*)calloc(2 * (size_t)nthread, sizeof(float));

#ifdef _OPENMP

    if (hostthreads <= 0) {
        hostthreads = omp_get_max_threads();
    }

#else
    hostthreads = 1;
#endif
    sink_t* sinks = (sink_t*)calloc((size_t)hostthreads, sizeof(sink_t));
    struct timespec ts0, ts1;
    clock_gettime(CLOCK_MONOTONIC, &ts0);
    #pragma omp parallel num_threads(hostthreads)
    {
#ifdef _OPENMP
        const int tid = omp_get_thread_num();
#else
        const int tid = 0;
#endif
        sink_t* s = sinks + tid;
        s->field = (float*)calloc(fieldlen * (rfforward ? 4 : 2), sizeof(float));

        if ((g.debuglevel & (2u | 8u)) && res->traj && g.maxjumpdebug) {
            s->traj = (float*)calloc((size_t)g.maxjumpdebug * 6, sizeof(float));
        }

        if (cfg->issavedet) {
            s->detp = (float*)calloc((size_t)g.maxdetphoton * (g.reclen ? g.reclen : 1), sizeof(float));

            if (cfg->issaveseed) {
                s->detseed = (uint64_t*)calloc((size_t)g.maxdetphoton * 2, sizeof(uint64_t));
            }
        }

        #pragma omp for schedule(dynamic, 16)

        for (long idx = 0; idx < (long)nthread; idx++) {
            work_item(&g, s, (int)idx, seeds, genergy, nthread);
        }
    }
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    res->runtime_ms = (double)(ts1.tv_sec - ts0.tv_sec) * 1e3 + (double)(ts1.tv_nsec - ts0.tv_nsec) * 1e-6;
    int rc = 0;

    /* fold the shadow half and sum host threads (src/mcx_host.cpp:1252-1258, 1292-1296) */
    if (res->field) {
        if (res->fieldlen < fieldlen * (rfforward ? 2 : 1)) {
            rc = -2;
        } else {
            for (size_t i = 0; i < fieldlen; i++) {
                float acc = 0.f, acc_im = 0.f;

                for (int t = 0; t < hostthreads; t++) {
                    float val = sinks[t].field[i];

                    if (!(g.debuglevel & 1u)) {
                        val += sinks[t].field[i + fieldlen];
                    }

                    acc += val;

                    if (rfforward) {        /* third + fourth quarter: imaginary primary + shadow (src/mcx_host.cpp:1263-1276) */
                        acc_im += sinks[t].field[i + 2 * fieldlen] + sinks[t].field[i + 3 * fieldlen];
                    }
                }

                res->field[i] = acc;

                if (rfforward) {
                    res->field[i + fieldlen] = acc_im;
                }
            }
        }
    }

    res->fieldlen = fieldlen * (rfforward ? 2 : 1);
    /* src/mcx_host.cpp:1303-1306 */
    double etot = 0.0, eesc = 0.0;

    for (size_t i = 0; i < nthread; i++) {
        eesc += genergy[i << 1];
        etot += genergy[(i << 1) + 1];
    }

    res->energytot = etot;
    res->energyesc = eesc;

    if (res->energy && !replay) {       /* the caller sized this buffer for ITS work-item count */
        memcpy(res->energy, genergy, sizeof(float) * 2 * nthread);
    }

    res->reclen = g.reclen;
    res->detected = 0;
    res->trajcount = 0;
    res->n_segment = res->n_deposit = res->n_scatter = 0;
    uint32_t saved = 0;

    for (int t = 0; t < hostthreads; t++) {
        sink_t* s = sinks + t;
        res->detected += s->detcount;
        const uint32_t n = s->detcount < g.maxdetphoton ? s->detcount : g.maxdetphoton;

        for (uint32_t k = 0; k < n && res->detphoton && saved < res->detcap; k++, saved++) {
            memcpy(res->detphoton + (size_t)saved * g.reclen, s->detp + (size_t)k * g.reclen, sizeof(float) * g.reclen);

            if (res->seeddata && cfg->issaveseed) {
                res->seeddata[2 * (size_t)saved] = s->detseed[2 * (size_t)k];
                res->seeddata[2 * (size_t)saved + 1] = s->detseed[2 * (size_t)k + 1];
            }
        }

        /* trajectory records of the host threads one after the other (oracle/ref_driver.cpp does the same) */
        const uint32_t ntraj = s->trajcount < g.maxjumpdebug ? s->trajcount : g.maxjumpdebug;

        for (uint32_t k = 0; s->traj && k < ntraj && res->trajcount < res->trajcap; k++, res->trajcount++) {
            memcpy(res->traj + (size_t)res->trajcount * 6, s->traj + (size_t)k * 6, sizeof(float) * 6);
        }

        res->n_segment += s->n_segment;
        res->n_deposit += s->n_atomic;
        res->n_scatter += s->n_log;
        free(s->field);
        free(s->traj);
        free(s->detp);
        free(s->detseed);
    }

    res->n_launch = (uint64_t)llround(etot);
    free(sinks);
    free(genergy);
    free(seeds);
    free(sharedtab);
    free(gproperty);
    return rc;
}

/* ------------------------------------------------------------------- unit-level hooks (same API as _ref) */

int mcxo_rng(const uint32_t* seeds, uint32_t n, uint32_t ndraw, float* out, uint64_t* state_out) {
    for (uint32_t i = 0; i < n; i++) {
        uint64_t t[2];
        rng_init(t, seeds + (size_t)i * 4);

        for (uint32_t k = 0; k < ndraw; k++) {
            out[(size_t)i * ndraw + k] = rand_uniform01(t);
        }

        if (state_out) {
            state_out[2 * i] = t[0];
            state_out[2 * i + 1] = t[1];
        }
    }

    return 0;
}

/* the stepping statements of :2677-2680, 2708-2747 with an unlimited scattering length */
int mcxo_trace(const mcxb_f4* p0, const mcxb_f4* v0, uint32_t n, uint32_t nstep,
               uint32_t dimx, uint32_t dimy, uint32_t dimz, float musp, mcxb_trace_step* out) {
    for (uint32_t i = 0; i < n; i++) {
        f4 p = {p0[i].x, p0[i].y, p0[i].z, p0[i].w}, v = {v0[i].x, v0[i].y, v0[i].z, v0[i].w};
        s4 id = {(short)floorf(p.x), (short)floorf(p.y), (short)floorf(p.z), -1};

        for (uint32_t k = 0; k < nstep; k++) {
            mcxb_trace_step* o = out + (size_t)i * nstep + k;
            const float dist = hitgrid(NULL, &p, &v, &id);
            const float slen = dist * musp;
            const float fz = slen / musp;
            p.x = p.x + fz * v.x;
            p.y = p.y + fz * v.y;
            p.z = p.z + fz * v.z;

            if (id.w == 0) {
                id.x += (v.x > 0.f ? 1 : -1);
            }

            if (id.w == 1) {
                id.y += (v.y > 0.f ? 1 : -1);
            }

            if (id.w == 2) {
                id.z += (v.z > 0.f ? 1 : -1);
            }

            o->dist = fz;
            o->px = p.x;
            o->py = p.y;
            o->pz = p.z;
            o->ix = id.x;
            o->iy = id.y;
            o->iz = id.z;
            o->face = id.w;

            if ((unsigned short)id.x >= dimx || (unsigned short)id.y >= dimy || (unsigned short)id.z >= dimz) {
                o->idx1d = (id.x < 0 || id.y < 0 || id.z < 0) ? OUTSIDE_VOLUME_MIN : OUTSIDE_VOLUME_MAX;

                for (uint32_t j = k + 1; j < nstep; j++) {
                    out[(size_t)i * nstep + j] = *o;
                }

                break;
            }

            o->idx1d = (uint32_t)(id.z * (int)(dimx * dimy) + id.y * (int)dimx + id.x);
        }
    }

    return 0;
}

int mcxo_scalar(const float* a, const int32_t* dir, uint32_t n, float* nextafter_out,
                const mcxb_f4* v, const float* n1, const float* n2, const int32_t* face, uint32_t m, float* rcoef_out) {
    for (uint32_t i = 0; i < n; i++) {
        nextafter_out[i] = mcx_nextafterf(a[i], dir[i]);
    }

    for (uint32_t i = 0; i < m; i++) {
        const f4 vv = {v[i].x, v[i].y, v[i].z, v[i].w};
        rcoef_out[i] = reflectcoeff(&vv, n1[i], n2[i], (short)face[i]);
    }

    return 0;
}

int mcxo_rotate(mcxb_f4* v, const float* stheta, const float* ctheta, const float* sphi, const float* cphi, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        f4 vv = {v[i].x, v[i].y, v[i].z, v[i].w};
        rotatevector(&vv, stheta[i], ctheta[i], sphi[i], cphi[i]);
        v[i].x = vv.x;
        v[i].y = vv.y;
        v[i].z = vv.z;
        v[i].w = vv.w;
    }

    return 0;
}

int mcxo_transmit(mcxb_f4* v, const float* n1, const float* n2, const int32_t* face, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        f4 vv = {v[i].x, v[i].y, v[i].z, v[i].w};
        transmit(&vv, n1[i], n2[i], (short)face[i]);
        v[i].x = vv.x;
        v[i].y = vv.y;
        v[i].z = vv.z;
        v[i].w = vv.w;
    }

    return 0;
}

unsigned int mcxo_configsize(void) {
    return (unsigned int)sizeof(mcxb_config);
}
