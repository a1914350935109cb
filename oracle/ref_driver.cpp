/*
 * ref_driver.cpp -- host driver around the REFERENCE kernel compiled through clshim.h
 * (TEST INFRASTRUCTURE ONLY; product code never links or loads this).
 *
 * It does what the reference host does around clEnqueueNDRangeKernel, restated for a CPU run:
 *   - packs MCXParam the way mcx_run_simulation does (reference src/mcx_host.cpp:494-524, 674-694,
 *     1076, 1095-1104),
 *   - seeds every work-item from ONE glibc srand()/rand() stream (:696-700, 759-768),
 *   - calls the kernel once per work-item with get_local_size()==1 and a private zeroed __local
 *     scratch,
 *   - folds the shadow half of the field (:1252-1258) and sums the per-thread energies (:1303-1306).
 * Work-items are spread over host threads with OpenMP; every host thread owns a private field and
 * detected-photon buffer which are summed / concatenated afterwards.
 */
#include "clshim.h"
#include "oracle_api.h"
#include <vector>
#include <chrono>
#include <omp.h>

thread_local ClShimWorkItem clshim_wi = {0, 0, 1, 1, 0};
thread_local unsigned long long clshim_cnt_isgreater = 0, clshim_cnt_xchg = 0, clshim_cnt_log = 0;

namespace refd {
#include "mcx_core_patched.cl"
}

typedef void (*ref_kernel_fn)(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*,
                              const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*,
                              void*, const void*, float*, float*, int*);

thread_local unsigned int* mcxref_jumpdebug = NULL;
thread_local float* mcxref_debugdata = NULL;
thread_local const void* mcxref_smatrix = NULL;      /* Mueller-matrix tables of a polarised run (gsmatrix, kernel argument 19) */

#define REF_DECL(name) \
    extern "C" void mcxref_kernel_##name##_r0_d0(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*, const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*, void*, const void*, float*, float*, int*); \
    extern "C" void mcxref_kernel_##name##_r1_d0(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*, const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*, void*, const void*, float*, float*, int*); \
    extern "C" void mcxref_kernel_##name##_r0_d1(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*, const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*, void*, const void*, float*, float*, int*); \
    extern "C" void mcxref_kernel_##name##_r1_d1(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*, const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*, void*, const void*, float*, float*, int*);
#define REF_ROW(name) { mcxref_kernel_##name##_r0_d0, mcxref_kernel_##name##_r1_d0, mcxref_kernel_##name##_r0_d1, mcxref_kernel_##name##_r1_d1 }

REF_DECL(pencil) REF_DECL(isotropic) REF_DECL(cone) REF_DECL(gaussian) REF_DECL(planar) REF_DECL(pattern)
REF_DECL(fourier) REF_DECL(arcsine) REF_DECL(disk) REF_DECL(fourierx) REF_DECL(fourierx2d) REF_DECL(zgaussian)
REF_DECL(line) REF_DECL(slit) REF_DECL(pencilarray) REF_DECL(pattern3d) REF_DECL(hyperboloid) REF_DECL(ring)

/* indexed by srctype (reference src/mcx_const.h:75-92), then [reflect + 2*savedet] */
static const ref_kernel_fn ref_kernels[18][4] = {
    REF_ROW(pencil), REF_ROW(isotropic), REF_ROW(cone), REF_ROW(gaussian), REF_ROW(planar), REF_ROW(pattern),
    REF_ROW(fourier), REF_ROW(arcsine), REF_ROW(disk), REF_ROW(fourierx), REF_ROW(fourierx2d), REF_ROW(zgaussian),
    REF_ROW(line), REF_ROW(slit), REF_ROW(pencilarray), REF_ROW(pattern3d), REF_ROW(hyperboloid), REF_ROW(ring)
};

/* continuous-media and split-voxel builds of the pencil-source kernel (-DMED_TYPE=99..104, 97), [format - 99 | 6][reflect + 2*savedet] */
#define REF_MDECL(m) \
    extern "C" void mcxref_kernel_pencil_m##m##_r0_d0(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*, const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*, void*, const void*, float*, float*, int*); \
    extern "C" void mcxref_kernel_pencil_m##m##_r1_d0(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*, const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*, void*, const void*, float*, float*, int*); \
    extern "C" void mcxref_kernel_pencil_m##m##_r0_d1(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*, const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*, void*, const void*, float*, float*, int*); \
    extern "C" void mcxref_kernel_pencil_m##m##_r1_d1(const unsigned int*, float*, float*, unsigned int*, float*, const void*, float*, const void*, volatile unsigned int*, unsigned int*, unsigned long*, float*, float*, void*, const void*, float*, float*, int*);
#define REF_MROW(m) { mcxref_kernel_pencil_m##m##_r0_d0, mcxref_kernel_pencil_m##m##_r1_d0, mcxref_kernel_pencil_m##m##_r0_d1, mcxref_kernel_pencil_m##m##_r1_d1 }
REF_MDECL(97) REF_MDECL(99) REF_MDECL(100) REF_MDECL(101) REF_MDECL(102) REF_MDECL(103) REF_MDECL(104)
static const ref_kernel_fn ref_media_kernels[7][4] = { REF_MROW(99), REF_MROW(100), REF_MROW(101), REF_MROW(102), REF_MROW(103), REF_MROW(104), REF_MROW(97) };

static inline float4 to_f4(const mcxb_f4& a) {
    return float4(a.x, a.y, a.z, a.w);
}

/* the reference's rule for compiling the reflection code in (src/mcx_host.cpp:945-956) */
static int ref_needs_reflection(const mcxb_config* cfg) {
    int allabsorb = 1, allunknown = 1;

    for (int i = 0; i < 6; i++) {
        if (cfg->bc[i] != MCXB_BC_ABSORB) {
            allabsorb = 0;
        }

        if (cfg->bc[i] != 0) {
            allunknown = 0;
        }
    }

    /* strcmp() in the reference stops at the first NUL, so an all-zero bc string equals "unknown" */
    if (cfg->bc[0] == 0) {
        allunknown = 1;
        allabsorb = 0;
    }

    return cfg->isreflect || (!allabsorb && !allunknown);
}

extern "C" void mcxref_seeds(int seed, uint64_t skip_records, uint64_t nrecords, uint32_t* out4) {
    srand(seed > 0 ? (unsigned)seed : 1u);

    for (uint64_t i = 0; i < skip_records * 4; i++) {
        (void)rand();
    }

    for (uint64_t i = 0; i < nrecords * 4; i++) {
        out4[i] = (uint32_t)rand();
    }
}

extern "C" int mcxref_run(const mcxb_config* cfg, uint32_t nthread, int hostthreads, mcxo_result* res) {
    using namespace refd;

    if (cfg->abi_version != MCXB_ABI_VERSION || cfg->srctype < 0 || cfg->srctype >= 18 || nthread == 0) {
        return -1;
    }

    const unsigned int dimxyz = cfg->dimx * cfg->dimy * cfg->dimz;
    const unsigned int maxgate = (unsigned int)((cfg->tend - cfg->tstart) / cfg->tstep + 0.5);
    const unsigned int nsrcvol = (cfg->srctype == MCXB_SRC_PATTERN || cfg->srctype == MCXB_SRC_PATTERN3D) ? cfg->srcnum
                                 : ((cfg->srcid < 0) ? (cfg->extrasrclen + 1) : 1);
    /* photon replay (src/mcx_host.cpp:684-689, 722-737): one volume per detector when replaydet == -1.  The
     * reference maps work-item t to record t*threadphoton + min(t, oddphoton-1) + k (src/mcx_core.cl:1591), which
     * is the identity only when every work-item owns at most one photon, so a replay runs with nphoton+1 work-items */
    const bool replay = cfg->replay_seed != NULL;
    const unsigned int nrepvol = (replay && cfg->replaydet == -1) ? (cfg->detnum ? cfg->detnum : 1) : 1;

    if (replay) {
        nthread = (uint32_t)cfg->nphoton + 1;
    }

    const size_t fieldlen = (size_t)dimxyz * maxgate * nsrcvol * nrepvol;
    /* RF outputs (src/mcx_host.cpp:473, 770, 1243-1276): a forward run with omega > 0 keeps real parts in [0,F) + shadow
     * [F,2F) and imaginary parts in [2F,3F) + shadow [3F,4F); the RF replay Jacobian keeps real parts in [0,F) and imaginary
     * parts in [F,2F).  The scattering-site form (otRFmus) writes its imaginary part at +2F like the forward run (:2601),
     * which the reference's host neither allocates nor reads; this driver allocates 4F always and folds it from there. */
    const bool rfforward = cfg->omega > 0.f && !replay;
    const bool rfreplay = replay && (cfg->outputtype == 6 || cfg->outputtype == 8) && cfg->omega > 0.f;
    const size_t planes = (rfforward || rfreplay) ? 2 : 1;

    const unsigned int flag = cfg->issavedet ? cfg->savedetflag : 0;
    const unsigned int partialdata = (cfg->medianum - 1) * (SAVE_NSCAT(flag) + SAVE_PPATH(flag) + SAVE_MOM(flag));
    const unsigned int w0offset = partialdata + 4;
    const unsigned int reclen = partialdata + SAVE_DETID(flag) + 3 * (SAVE_PEXIT(flag) + SAVE_VEXIT(flag)) + SAVE_W0(flag) + 4 * SAVE_IQUV(flag);
    unsigned int is2d = (cfg->dimx == 1 ? 1 : (cfg->dimy == 1 ? 2 : (cfg->dimz == 1 ? 3 : 0)));

    if (is2d) {
        is2d = is2d * (((cfg->dimx > 1) + (cfg->dimy > 1) + (cfg->dimz > 1)) == 2);
    }

    MCXParam param;
    memset(&param, 0, sizeof(param));
    param.src.pos = to_f4(cfg->src.pos);
    param.src.dir = to_f4(cfg->src.dir);
    param.src.param1 = to_f4(cfg->src.param1);
    param.src.param2 = to_f4(cfg->src.param2);
    param.extrasrclen = cfg->extrasrclen;
    param.srcid = cfg->srcid;
    param.maxidx = float4((float)cfg->dimx, (float)cfg->dimy, (float)cfg->dimz, 0.f);
    param.dimlen.x = cfg->dimx;
    param.dimlen.y = cfg->dimx * cfg->dimy;
    param.dimlen.z = dimxyz;
    param.dimlen.w = (unsigned int)fieldlen;
    param.minstep = cfg->unitinmm;          /* steps.{x,y,z} == unitinmm after mcx_preprocess */
    param.twin0 = cfg->tstart;
    param.twin1 = cfg->tstart + cfg->tstep * maxgate;
    param.tmax = cfg->tend;
    param.oneoverc0 = R_C0 * cfg->unitinmm;
    param.save2pt = (uint)cfg->issave2pt;
    param.doreflect = (uint)cfg->isreflect;
    param.savedet = (uint)cfg->issavedet;
    param.Rtstep = 1.f / cfg->tstep;
    param.minenergy = cfg->minenergy;
    param.minaccumtime = param.minstep * R_C0 * cfg->unitinmm;
    param.maxdetphoton = cfg->maxdetphoton;
    param.maxmedia = cfg->medianum - 1;
    param.detnum = cfg->detnum;
    param.voidtime = cfg->voidtime;
    param.srctype = cfg->srctype;
    param.maxvoidstep = (uint)cfg->maxvoidstep;
    param.issaveexit = (SAVE_PEXIT(flag) && SAVE_VEXIT(flag));
    param.issaveseed = cfg->issaveseed > 0;
    param.issaveref = (uint)cfg->issaveref;
    param.isspecular = cfg->isspecular > 0;
    param.maxgate = maxgate;
    param.seed = replay ? -999 /* SEED_FROM_FILE */ : cfg->seed;
    param.replaydet = cfg->replaydet;
    param.outputtype = (uint)cfg->outputtype;
    param.threadphoton = (uint)(cfg->nphoton / nthread);
    param.oddphoton = (int)(cfg->nphoton - (uint64_t)param.threadphoton * nthread);
    param.debuglevel = cfg->debuglevel & (1u | 2u | 8u);        /* MCX_DEBUG_RNG, MCX_DEBUG_MOVE, MCX_DEBUG_MOVE_ONLY */
    param.maxjumpdebug = cfg->maxjumpdebug;
    param.savedetflag = flag;
    param.reclen = reclen;
    param.partialdata = partialdata;
    param.w0offset = w0offset;
    param.mediaformat = cfg->mediaformat > 4 ? cfg->mediaformat : 1;
    param.gscatter = cfg->gscatter;
    param.is2d = is2d;
    param.srcnum = cfg->srcnum ? cfg->srcnum : 1;
    param.maxpolmedia = cfg->smatrix ? cfg->polmedianum : 0;     /* src/mcx_host.cpp:522-523 */
    param.istrajstokes = 0;
    param.s0 = to_f4(cfg->srciquv);
    param.omega = cfg->omega;
    /* inverse-CDF tables live at the head of the __local scratch (src/mcx_core.cl:2354-2383, src/mcx_host.cpp:1013) */
    param.nphase = cfg->invcdf ? cfg->nphase : 0;
    param.nphaselen = param.nphase + (param.nphase & 1);
    param.nangle = cfg->angleinvcdf ? cfg->nangle : 0;
    param.nanglelen = param.nangle + (param.nangle & 1);
    memcpy(param.bc, cfg->bc, 12);

    /* media table followed by the extra sources (src/mcx_host.cpp:746-751) */
    std::vector<float4> gproperty(cfg->medianum + 4 * cfg->extrasrclen);

    for (unsigned int i = 0; i < cfg->medianum; i++) {
        gproperty[i] = to_f4(cfg->prop[i]);
    }

    for (unsigned int i = 0; i < cfg->extrasrclen; i++) {
        gproperty[cfg->medianum + 4 * i + 0] = to_f4(cfg->srcdata[i].pos);
        gproperty[cfg->medianum + 4 * i + 1] = to_f4(cfg->srcdata[i].dir);
        gproperty[cfg->medianum + 4 * i + 2] = to_f4(cfg->srcdata[i].param1);
        gproperty[cfg->medianum + 4 * i + 3] = to_f4(cfg->srcdata[i].param2);
    }

    std::vector<float4> gdetpos(cfg->detnum ? cfg->detnum : 1);

    for (unsigned int i = 0; i < cfg->detnum; i++) {
        gdetpos[i] = to_f4(cfg->detpos[i]);
    }

    std::vector<uint32_t> seeds((size_t)nthread * 4 + 4);

    if (replay) {
        /* gseed holds the recorded RNG states (src/mcx_host.cpp:724); gpu_rng_init still reads one record per work-item */
        memset(seeds.data(), 0, seeds.size() * 4);
        memcpy(seeds.data(), cfg->replay_seed, 16 * (size_t)cfg->nphoton);
    } else {
        mcxref_seeds(cfg->seed, cfg->seed_skip, nthread, seeds.data());
    }

    ref_kernel_fn kern = ref_kernels[cfg->srctype][(ref_needs_reflection(cfg) ? 1 : 0) + (cfg->issavedet ? 2 : 0)];

    if (cfg->mediaformat > 4) {
        if (((cfg->mediaformat < 99 || cfg->mediaformat > 104) && cfg->mediaformat != 97) || cfg->srctype != 0) {
            return -3;      /* only the pencil-source builds of MED_TYPE 97 and 99..104 exist */
        }

        kern = ref_media_kernels[cfg->mediaformat == 97 ? 6 : cfg->mediaformat - 99][(ref_needs_reflection(cfg) ? 1 : 0) + (cfg->issavedet ? 2 : 0)];
    }


    if (hostthreads <= 0) {
        hostthreads = omp_get_max_threads();
    }

    const size_t rawlen = (param.debuglevel & 1u) ? fieldlen : fieldlen * 2;
    const unsigned int detcap = cfg->maxdetphoton;
    std::vector<std::vector<float> > tfield(hostthreads), tdet(hostthreads);
    std::vector<std::vector<unsigned long> > tseed(hostthreads);
    std::vector<unsigned int> tdetcount(hostthreads, 0);
    std::vector<unsigned long long> cnt_seg(hostthreads, 0), cnt_dep(hostthreads, 0), cnt_log(hostthreads, 0);
    std::vector<float> genergy((size_t)nthread * 2, 0.f);
    const size_t sharedbytes = 4 * (size_t)(param.nphaselen + param.nanglelen) + 4 * (size_t)(w0offset + param.srcnum + 2) + 16 * param.issaveseed + 64;

    const bool wanttraj = (param.debuglevel & (2u | 8u)) != 0 && res->traj != NULL;
    std::vector<std::vector<float> > ttraj(hostthreads);
    std::vector<unsigned int> ttrajcount(hostthreads, 0);

    auto t0 = std::chrono::steady_clock::now();
    #pragma omp parallel num_threads(hostthreads)
    {
        const int tid = omp_get_thread_num();
        tfield[tid].assign(fieldlen * 4, 0.f);
        mcxref_smatrix = cfg->smatrix;

        if (wanttraj) {
            /* every host thread records into its own buffer with its own counter (the kernel's atomic_inc, :930) */
            ttraj[tid].assign((size_t)param.maxjumpdebug * 6, 0.f);
            mcxref_jumpdebug = &ttrajcount[tid];
            mcxref_debugdata = ttraj[tid].data();
        } else {
            mcxref_jumpdebug = NULL;
            mcxref_debugdata = NULL;
        }

        if (cfg->issavedet) {
            tdet[tid].assign((size_t)detcap * (reclen ? reclen : 1), 0.f);

            if (cfg->issaveseed) {
                tseed[tid].assign((size_t)detcap * 2, 0ul);
            }
        }

        std::vector<unsigned long> shared((sharedbytes + 7) / 8);
        unsigned int progress = 0;
        clshim_cnt_isgreater = clshim_cnt_xchg = clshim_cnt_log = 0;

        #pragma omp for schedule(dynamic, 16)

        for (long idx = 0; idx < (long)nthread; idx++) {
            clshim_wi.global_id = (size_t)idx;
            clshim_wi.local_id = 0;
            clshim_wi.local_size = 1;
            clshim_wi.num_groups = nthread;
            clshim_wi.group_id = (size_t)idx;
            std::fill(shared.begin(), shared.end(), 0ul);
            kern(cfg->vol, tfield[tid].data(), genergy.data(), seeds.data(),
                 cfg->issavedet ? tdet[tid].data() : NULL, gproperty.data(), (float*)cfg->srcpattern,
                 gdetpos.data(), &progress, &tdetcount[tid],
                 cfg->issaveseed ? tseed[tid].data() : NULL, (float*)cfg->invcdf, (float*)cfg->angleinvcdf, shared.data(), &param,
                 (float*)cfg->replay_weight, (float*)cfg->replay_tof, (int*)cfg->replay_detid);
        }

        cnt_seg[tid] = clshim_cnt_isgreater;
        cnt_dep[tid] = clshim_cnt_xchg / 2;
        cnt_log[tid] = clshim_cnt_log;
    }
    auto t1 = std::chrono::steady_clock::now();
    res->runtime_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    (void)rawlen;

    /* fold shadow half and sum host threads (src/mcx_host.cpp:1252-1258, 1292-1296) */
    if (res->field) {
        if (res->fieldlen < fieldlen * planes) {
            return -2;
        }

        for (size_t i = 0; i < fieldlen; i++) {
            float acc = 0.f, acc_im = 0.f;

            for (int t = 0; t < hostthreads; t++) {
                float v = tfield[t][i];

                if (!(param.debuglevel & 1u) && !(rfreplay && cfg->outputtype == 6)) {
                    v += tfield[t][i + fieldlen];
                }

                acc += v;

                if (rfforward || (rfreplay && cfg->outputtype == 8)) {
                    acc_im += tfield[t][i + 2 * fieldlen] + tfield[t][i + 3 * fieldlen];
                } else if (rfreplay) {
                    acc_im += tfield[t][i + fieldlen];
                }
            }

            res->field[i] = acc;

            if (planes == 2) {
                res->field[i + fieldlen] = acc_im;
            }
        }
    }

    res->fieldlen = fieldlen * planes;

    double etot = 0.0, eesc = 0.0;

    for (size_t i = 0; i < nthread; i++) {
        eesc += genergy[i << 1];
        etot += genergy[(i << 1) + 1];
    }

    res->energytot = etot;
    res->energyesc = eesc;

    if (res->energy && !replay) {       /* the caller sized this buffer for ITS work-item count */
        memcpy(res->energy, genergy.data(), sizeof(float) * 2 * nthread);
    }

    res->reclen = reclen;
    res->detected = 0;
    unsigned int saved = 0;

    for (int t = 0; t < hostthreads; t++) {
        res->detected += tdetcount[t];
        unsigned int n = std::min(tdetcount[t], detcap);

        for (unsigned int k = 0; k < n && res->detphoton && saved < res->detcap; k++, saved++) {
            memcpy(res->detphoton + (size_t)saved * reclen, tdet[t].data() + (size_t)k * reclen, sizeof(float) * reclen);

            if (res->seeddata && cfg->issaveseed) {
                res->seeddata[2 * (size_t)saved] = tseed[t][2 * (size_t)k];
                res->seeddata[2 * (size_t)saved + 1] = tseed[t][2 * (size_t)k + 1];
            }
        }
    }

    res->trajcount = 0;

    for (int t = 0; wanttraj && t < hostthreads; t++) {
        const unsigned int n = std::min(ttrajcount[t], param.maxjumpdebug);

        for (unsigned int k = 0; k < n && res->trajcount < res->trajcap; k++, res->trajcount++) {
            memcpy(res->traj + (size_t)res->trajcount * 6, ttraj[t].data() + (size_t)k * 6, sizeof(float) * 6);
        }
    }

    res->n_segment = res->n_deposit = res->n_scatter = 0;

    for (int t = 0; t < hostthreads; t++) {
        res->n_segment += cnt_seg[t];
        res->n_deposit += cnt_dep[t];
        res->n_scatter += cnt_log[t];
    }

    res->n_launch = (uint64_t)llround(etot);
    return 0;
}

/* ---- unit-level known-answer hooks: they call the reference's own helper functions ---------- */

extern "C" int mcxref_rng(const uint32_t* seeds, uint32_t n, uint32_t ndraw, float* out, uint64_t* state_out) {
    using namespace refd;

    for (uint32_t i = 0; i < n; i++) {
        RandType t[RAND_BUF_LEN];
        gpu_rng_init(t, (uint*)seeds, (int)i);

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

extern "C" int mcxref_trace(const mcxb_f4* p0, const mcxb_f4* v0, uint32_t n, uint32_t nstep,
                            uint32_t dimx, uint32_t dimy, uint32_t dimz, float musp, mcxb_trace_step* out) {
    using namespace refd;

    for (uint32_t i = 0; i < n; i++) {
        float4 p = to_f4(p0[i]), v = to_f4(v0[i]);
        short4 flipdir((short)floorf(p.x), (short)floorf(p.y), (short)floorf(p.z), -1);

        for (uint32_t k = 0; k < nstep; k++) {
            mcxb_trace_step* o = out + (size_t)i * nstep + k;
            /* the stepping statements of src/mcx_core.cl:2677-2680, 2708-2747 with f.x = +inf */
            float dist = hitgrid(&p, &v, &flipdir);
            float slen = dist * musp;
            float fz = native_divide(slen, musp);
            p.x = p.x + fz * v.x;
            p.y = p.y + fz * v.y;
            p.z = p.z + fz * v.z;

            if (flipdir.w == 0) {
                flipdir.x += (v.x > 0.f ? 1 : -1);
            }

            if (flipdir.w == 1) {
                flipdir.y += (v.y > 0.f ? 1 : -1);
            }

            if (flipdir.w == 2) {
                flipdir.z += (v.z > 0.f ? 1 : -1);
            }

            o->dist = fz;
            o->px = p.x;
            o->py = p.y;
            o->pz = p.z;
            o->ix = flipdir.x;
            o->iy = flipdir.y;
            o->iz = flipdir.z;
            o->face = flipdir.w;

            if ((ushort)flipdir.x >= dimx || (ushort)flipdir.y >= dimy || (ushort)flipdir.z >= dimz) {
                o->idx1d = (flipdir.x < 0 || flipdir.y < 0 || flipdir.z < 0) ? OUTSIDE_VOLUME_MIN : OUTSIDE_VOLUME_MAX;

                for (uint32_t j = k + 1; j < nstep; j++) {
                    out[(size_t)i * nstep + j] = *o;
                }

                break;
            }

            o->idx1d = (uint32_t)(flipdir.z * (int)(dimx * dimy) + flipdir.y * (int)dimx + flipdir.x);
        }
    }

    return 0;
}

extern "C" int mcxref_scalar(const float* a, const int32_t* dir, uint32_t n, float* nextafter_out,
                             const mcxb_f4* v, const float* n1, const float* n2, const int32_t* face, uint32_t m, float* rcoef_out) {
    using namespace refd;

    for (uint32_t i = 0; i < n; i++) {
        nextafter_out[i] = mcx_nextafterf(a[i], dir[i]);
    }

    for (uint32_t i = 0; i < m; i++) {
        float4 vv = to_f4(v[i]);
        rcoef_out[i] = reflectcoeff(&vv, n1[i], n2[i], (short)face[i]);
    }

    return 0;
}

/* rotatevector (src/mcx_core.cl:1025-1042) and transmit (:1044-1055) */
extern "C" int mcxref_rotate(mcxb_f4* v, const float* stheta, const float* ctheta, const float* sphi, const float* cphi, uint32_t n) {
    using namespace refd;

    for (uint32_t i = 0; i < n; i++) {
        float4 vv = to_f4(v[i]);
        rotatevector(&vv, stheta[i], ctheta[i], sphi[i], cphi[i]);
        v[i].x = vv.x;
        v[i].y = vv.y;
        v[i].z = vv.z;
        v[i].w = vv.w;
    }

    return 0;
}

extern "C" int mcxref_transmit(mcxb_f4* v, const float* n1, const float* n2, const int32_t* face, uint32_t n) {
    using namespace refd;

    for (uint32_t i = 0; i < n; i++) {
        float4 vv = to_f4(v[i]);
        transmit(&vv, n1[i], n2[i], (short)face[i]);
        v[i].x = vv.x;
        v[i].y = vv.y;
        v[i].z = vv.z;
        v[i].w = vv.w;
    }

    return 0;
}

extern "C" unsigned int mcxref_configsize(void) {
    return (unsigned int)sizeof(mcxb_config);
}

extern "C" unsigned int mcxref_paramsize(void) {
    return (unsigned int)sizeof(refd::MCXParam);
}
