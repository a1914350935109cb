/*
 * engine.cu -- CUDA-runtime host module behind the C ABI of include/mcxb200.h.
 *
 * It replaces, for the photon-transport path only, what the reference's OpenCL host does inside
 * mcx_run_simulation (src/mcx_host.cpp:438-1849): parameter packing (:494-524, 674-694), thread/block
 * autoconfiguration (:586-633), per-thread seeding from one glibc rand() stream (:696-700, 759-768),
 * buffer upload (:739-830), kernel launch and timing window (:1078-1168), readback, accumulation
 * (:1242-1306) and normalisation (:1382-1465); and device enumeration (mcx_list_gpu, :252-432).
 *
 * There is no CPU path: every entry point that needs a device fails with a CUDA error code when no
 * device is present.
 */
#include "../../include/mcxb200.h"
#include "kernel_registry.h"

#include <cstdio>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <ctime>
#include <string>
#include <chrono>
#include <vector>
#include <thread>
#include <map>
#include <mutex>
#include <algorithm>

using namespace mcxb;

/* ------------------------------------------------------------------------------------------------- */
static thread_local std::string g_last_error;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_TRY(call)                                                                                         \
    do {                                                                                                     \
        cudaError_t e__ = (call);                                                                            \
        if (e__ != cudaSuccess) {                                                                            \
            return fail(MCXB_ERR_CUDA_BASE - (int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                                 \
        }                                                                                                    \
    } while (0)

extern "C" const char* mcxb_last_error(void) {
    return g_last_error.c_str();
}

/* not part of the public header: lets the other translation units of the library (engine_multi.cu) report through the
 * same per-thread message */
extern "C" void mcxb_set_last_error(const char* msg) {
    g_last_error = msg ? msg : "";
}

/* -------------------------------------------------------------------------------------------------
 * glibc-compatible rand() stream (TYPE_3 additive feedback generator, degree 31, separation 3), so the
 * per-thread seeds are the ones the reference host produces with srand(seed); rand() without
 * depending on the C library's hidden global state.
 * ------------------------------------------------------------------------------------------------- */
namespace {
struct GlibcRand {
    uint32_t r[31];
    int f, b;
    explicit GlibcRand(uint32_t seed) {
        if (seed == 0) {
            seed = 1;
        }

        int32_t word = (int32_t)seed;
        r[0] = (uint32_t)word;

        for (int i = 1; i < 31; i++) {
            const long hi = word / 127773, lo = word % 127773;
            long w = 16807 * lo - 2836 * hi;

            if (w < 0) {
                w += 2147483647;
            }

            word = (int32_t)w;
            r[i] = (uint32_t)word;
        }

        f = 3;
        b = 0;

        for (int i = 0; i < 310; i++) {
            next();
        }
    }
    uint32_t next() {
        r[f] += r[b];
        const uint32_t out = r[f] >> 1;
        f = (f + 1 == 31) ? 0 : f + 1;
        b = (b + 1 == 31) ? 0 : b + 1;
        return out;
    }
};
} // namespace

extern "C" void mcxb_fill_seeds(int32_t seed, uint64_t skip_records, uint64_t nrecords, uint32_t* out4) {
    GlibcRand g(seed > 0 ? (uint32_t)seed : (uint32_t)time(NULL));

    for (uint64_t i = 0; i < skip_records * 4; i++) {
        (void)g.next();
    }

    for (uint64_t i = 0; i < nrecords * 4; i++) {
        out4[i] = g.next();
    }
}

/* -------------------------------------------------------------------------------------------------
 * Buffer pool.  cudaFree / cudaFreeHost synchronise the device and unpin pages: measured on a B200 box the
 * teardown of one simulation cost 12 ms to more than a second, several times the rest of the host work of a call.
 * Front-ends call mcx_run_simulation in loops (wavelengths, sources, repetitions), always with the same buffer
 * sizes, so released buffers are kept by exact size per device and handed out again.  Cached bytes are bounded;
 * mcxb_release_cached_buffers() returns everything to the driver.
 * ------------------------------------------------------------------------------------------------- */
namespace {
struct BufferPool {
    std::mutex mu;
    std::multimap<std::pair<int, size_t>, void*> freedev, freehost;     /* (device, bytes) -> pointer; host: device = -1 */
    std::map<void*, std::pair<int, size_t> > live;
    size_t cacheddev = 0, cachedhost = 0;
    static constexpr size_t kMaxDev = (size_t)16 << 30, kMaxHost = (size_t)2 << 30;

    cudaError_t get(void** out, int device, size_t bytes, bool host) {
        bytes = std::max<size_t>(bytes, 16);
        std::lock_guard<std::mutex> lk(mu);
        auto& fl = host ? freehost : freedev;
        auto it = fl.find(std::make_pair(host ? -1 : device, bytes));

        if (it != fl.end()) {
            *out = it->second;
            fl.erase(it);
            (host ? cachedhost : cacheddev) -= bytes;
        } else {
            cudaError_t e = host ? cudaMallocHost(out, bytes) : cudaMalloc(out, bytes);

            if (e != cudaSuccess) {
                /* make room and try once more */
                cudaGetLastError();
                int cur = 0;
                cudaGetDevice(&cur);
                drop_locked();              /* frees on every device it holds buffers of */
                cudaSetDevice(cur);
                e = host ? cudaMallocHost(out, bytes) : cudaMalloc(out, bytes);

                if (e != cudaSuccess) {
                    *out = nullptr;
                    return e;
                }
            }
        }

        live[*out] = std::make_pair(host ? -1 : device, bytes);
        return cudaSuccess;
    }
    void put(void* ptr) {
        if (!ptr) {
            return;
        }

        std::lock_guard<std::mutex> lk(mu);
        auto it = live.find(ptr);

        if (it == live.end()) {
            return;
        }

        const std::pair<int, size_t> key = it->second;
        live.erase(it);
        const bool host = key.first < 0;
        size_t& cached = host ? cachedhost : cacheddev;

        if (cached + key.second > (host ? kMaxHost : kMaxDev)) {
            if (host) {
                cudaFreeHost(ptr);
            } else {
                cudaSetDevice(key.first);
                cudaFree(ptr);
            }

            return;
        }

        cached += key.second;
        (host ? freehost : freedev).insert(std::make_pair(key, ptr));
    }
    void drop_locked() {
        for (auto& kv : freedev) {
            cudaSetDevice(kv.first.first);
            cudaFree(kv.second);
        }

        for (auto& kv : freehost) {
            cudaFreeHost(kv.second);
        }

        freedev.clear();
        freehost.clear();
        cacheddev = cachedhost = 0;
    }
    void drop() {
        std::lock_guard<std::mutex> lk(mu);
        drop_locked();
    }
};
BufferPool& pool() {
    static BufferPool* p = new BufferPool();      /* intentionally leaked: no CUDA calls during static destruction */
    return *p;
}
template <typename T> cudaError_t dev_alloc(T** out, int device, size_t bytes) {
    return pool().get(reinterpret_cast<void**>(out), device, bytes, false);
}
template <typename T> cudaError_t host_alloc(T** out, size_t bytes) {
    return pool().get(reinterpret_cast<void**>(out), -1, bytes, true);
}
} // namespace

extern "C" void mcxb_release_cached_buffers(void) {
    pool().drop();
}

/* cudaGetDeviceProperties costs milliseconds (measured: up to 200 ms on a busy host) and the answer never changes:
 * one query per device per process */
static cudaError_t device_properties(int device, const cudaDeviceProp** out) {
    static std::mutex mu;
    static std::map<int, cudaDeviceProp> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(device);

    if (it == cache.end()) {
        cudaDeviceProp p;
        const cudaError_t e = cudaGetDeviceProperties(&p, device);

        if (e != cudaSuccess) {
            return e;
        }

        it = cache.insert(std::make_pair(device, p)).first;
    }

    *out = &it->second;
    return cudaSuccess;
}

/* the seed table of the last call: front-ends repeat runs with the same seed and launch shape */
static void cached_seeds(int32_t seed, uint64_t skip, uint64_t nrecords, std::vector<uint32_t>& out) {
    static std::mutex mu;
    static int32_t cseed = 0;
    static uint64_t cskip = 0;
    static std::vector<uint32_t> cache;
    std::lock_guard<std::mutex> lk(mu);

    if (seed <= 0 || seed != cseed || skip != cskip || cache.size() != nrecords * 4) {
        cache.resize(nrecords * 4);
        mcxb_fill_seeds(seed, skip, nrecords, cache.data());
        cseed = seed;
        cskip = skip;
    }

    out = cache;
}

/* ------------------------------------------------------------------------------------------------- */
static int cores_per_sm(int major, int minor) {
    /* same table idea as mcx_nv_corecount (src/mcx_host.cpp:224-240), extended to Hopper/Blackwell */
    if (major < 2) {
        return 8;
    }

    if (major == 2) {
        return minor == 0 ? 32 : 48;
    }

    if (major == 3) {
        return 192;
    }

    if (major == 5) {
        return 128;
    }

    if (major == 6) {
        return minor == 0 ? 64 : 128;
    }

    if (major == 7) {
        return 64;
    }

    if (major == 8) {
        return minor == 0 ? 64 : 128;
    }

    return 128;
}

extern "C" int mcxb_list_gpu(mcxb_gpuinfo* info, int maxinfo) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);

    if (e != cudaSuccess) {
        if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) {
            cudaGetLastError();
            return 0;
        }

        return fail(MCXB_ERR_CUDA_BASE - (int)e, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
    }

    for (int i = 0; i < n && i < maxinfo && info; i++) {
        const cudaDeviceProp* pp = nullptr;
        CU_TRY(device_properties(i, &pp));
        const cudaDeviceProp& p = *pp;
        mcxb_gpuinfo* g = info + i;
        memset(g, 0, sizeof(*g));
        strncpy(g->name, p.name, sizeof(g->name) - 1);
        g->id = i + 1;
        g->devcount = n;
        g->major = p.major;
        g->minor = p.minor;
        g->globalmem = p.totalGlobalMem;
        g->constmem = p.totalConstMem;
        g->sharedmem = p.sharedMemPerBlock;
        g->regcount = p.regsPerBlock;
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, i);
        g->clock_khz = khz;
        g->sm = p.multiProcessorCount;
        g->core = p.multiProcessorCount * cores_per_sm(p.major, p.minor);
        g->maxmpthread = p.maxThreadsPerMultiProcessor;
        g->autoblock = kBlock;
        g->autothread = (uint64_t)kBlock * MCXB_MINBLOCKS * p.multiProcessorCount;
        g->l2cache = (uint64_t)p.l2CacheSize;
    }

    return n;
}

/* ------------------------------------------------------------------------------------------------- */
struct mcxb_sim {
    int device = 0;
    mcxb_config cfg;                 /* scalar fields only; host pointers are not kept */
    SimParam P;
    PhotonKernelFn fn = nullptr;
    const char* kname = "";
    bool acc64 = true, media16 = false, media32 = false, rngdebug = false;
    uint32_t nblock = 0, nthread = 0;
    size_t smem = 0;
    uint64_t fieldlen = 0;
    uint32_t reclen = 0, maxgate = 0, nsrcvol = 1;
    /* device buffers */
    void* d_media = nullptr;
    void* d_field = nullptr;
    float* d_field32 = nullptr;
    float4* d_tables = nullptr;
    uint32_t* d_seeds = nullptr;
    float* d_det = nullptr;
    uint32_t* d_detcount = nullptr;
    unsigned long long* d_seedout = nullptr;
    unsigned long long* d_counter = nullptr;
    double* d_energy = nullptr;
    float* d_pattern = nullptr;
    float* d_invcdf = nullptr;
    unsigned long long* d_stats = nullptr;
    float* d_traj = nullptr;                     /* trajectory records (-D M) and their counter */
    uint32_t* d_trajcount = nullptr;
    uint64_t launched_photons = 0;               /* photons of the batches launched since the last reset (respin) */
    unsigned long long* d_rseed = nullptr;      /* replay inputs */
    float* d_rweight = nullptr;
    float* d_rtof = nullptr;
    int32_t* d_rdetid = nullptr;
    std::vector<float> h_srcpw;                  /* photon sharing: sum of each pattern (src/mcx_host.cpp:1357-1362) */
    std::vector<float> h_rweight;                /* host copies for the replay normalisation */
    std::vector<int32_t> h_rdetid;
    uint32_t nrepvol = 1;
    uint32_t rfplanes = 1;           /* 2 = real + imaginary volume sets (RF outputs) */
    uint64_t planelen = 0;           /* elements of one volume set: fieldlen = planelen * rfplanes */
    bool ext = false;                /* extended-physics kernel (polarised / RF) */
    float4* d_smatrix = nullptr;
    uint32_t acccopies = 1;                     /* replicated accumulator volumes (photon_kernel.cuh) */
    unsigned long long* h_progress = nullptr;   /* pinned word the progress poll copies the photon counter into */
    cudaStream_t pollstream = nullptr;
    /* pinned staging */
    float* h_field = nullptr;
    double* h_small = nullptr;       /* energy[2], detcount, stats[3] */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool finalized = false, launched = false;
    uint64_t launches = 0;
    float last_ms = 0.f;
};

/* det: 0 = no detector capture, 1 = the default record, 2 = any record flags (generic kernels take 1 and 2 alike) */
static const KernelEntry* find_kernel(int src, bool refl, int det, int mediabits, bool acc64, bool stats, bool common, bool queue = false, bool ext = false,
                                      bool bcodes = false) {
    typedef const KernelEntry* (*GroupFn)(int*);
    static const GroupFn groups[kNumGroups] = { mcxb_kernel_group_0, mcxb_kernel_group_1, mcxb_kernel_group_2, mcxb_kernel_group_3,
                                                mcxb_kernel_group_4, mcxb_kernel_group_5, mcxb_kernel_group_6, mcxb_kernel_group_7,
                                                mcxb_kernel_group_8, mcxb_kernel_group_9, mcxb_kernel_group_10, mcxb_kernel_group_11
                                              };
    const bool m16 = mediabits == 16, m32 = mediabits == 32;
    /* most specialised first: {source, common} -> {any source, common} -> {any source, generic} */
    const int wantsrc[3] = { src, (int)srcAny, (int)srcAny };
    const bool wantgen[3] = { false, false, true };

    for (int pass = common ? 0 : 2; pass < 3; pass++) {
        for (int g = 0; g < kNumGroups; g++) {
            int n = 0;
            const KernelEntry* e = groups[g](&n);

            for (int i = 0; i < n; i++) {
                const bool detok = wantgen[pass] ? ((e[i].savedet != 0) == (det != 0)) : (e[i].savedet == det);

                if (e[i].src == wantsrc[pass] && e[i].generic == wantgen[pass] && e[i].reflect == refl && detok &&
                        e[i].media16 == m16 && e[i].media32 == m32 && e[i].acc64 == acc64 && e[i].stats == stats && (e[i].queue != 0) == (queue && !wantgen[pass]) &&
                        e[i].ext == ext && e[i].bcodes == (bcodes && !wantgen[pass])) {
                    return e + i;
                }
            }
        }
    }

    return nullptr;
}

/* true when the configuration fits the compile-time assumptions of the GEN=false kernels (photon_kernel.cuh) */
static bool is_common_config(const mcxb_config* cfg, bool savedet, uint32_t nphase) {
    const bool is3d = cfg->dimx > 1 && cfg->dimy > 1 && cfg->dimz > 1;
    return is3d && nphase <= 2 && cfg->gscatter >= 1000000000u && cfg->issaveref == 0 && !(cfg->debuglevel & (MCXB_DEBUG_MOVE | MCXB_DEBUG_MOVE_ONLY)) &&
           cfg->issave2pt != 0 && cfg->replay_seed == nullptr && cfg->srcnum <= 1 &&
           (cfg->outputtype == MCXB_OT_FLUX || cfg->outputtype == MCXB_OT_FLUENCE || cfg->outputtype == MCXB_OT_ENERGY || cfg->outputtype == MCXB_OT_L ||
            (cfg->outputtype >= MCXB_OT_ADJOINT && cfg->outputtype <= MCXB_OT_ADJOINT_MUA_MUSP));      /* adjoint types run as fluence; energy / length: BCODES kernels */
}

/* the reference's rule for compiling the reflection code in (src/mcx_host.cpp:945-956) */
static bool needs_reflection(const mcxb_config* cfg) {
    bool allabsorb = true, allunknown = true;

    for (int i = 0; i < 6; i++) {
        if (cfg->bc[i] != MCXB_BC_ABSORB) {
            allabsorb = false;
        }

        if (cfg->bc[i] != MCXB_BC_UNKNOWN) {
            allunknown = false;
        }
    }

    if (cfg->bc[0] == 0) {      /* the reference compares C strings: a leading NUL reads as "unknown" */
        allunknown = true;
        allabsorb = false;
    }

    return cfg->isreflect || (!allabsorb && !allunknown);
}

/* the scattering-queue kernels are chosen when the mean mus per voxel is at most this (see sim_create_impl) */
static constexpr double kQueueMaxMus = 2.25;    /* measured crossover on cube60b with mus swept: gain 1.17 at 0.25, 1.08 at 1, 1.02 at 2, 1.00 at 2.5, 0.93 at 10 (profiles/r2_queue_crossover.log) */

static uint32_t count_gates(const mcxb_config* cfg) {
    return (uint32_t)((cfg->tend - cfg->tstart) / cfg->tstep + 0.5);
}

static float normalizer_impl(const mcxb_config* cfg, double energytot, bool replay) {
    /* src/mcx_host.cpp:1389-1396, 1449-1451; Vvox = steps.x*steps.y*steps.z with steps == unitinmm */
    float scale = 1.f;
    const float Vvox = cfg->unitinmm * cfg->unitinmm * cfg->unitinmm;
    const bool adjoint = cfg->outputtype >= MCXB_OT_ADJOINT && cfg->outputtype <= MCXB_OT_ADJOINT_MUA_MUSP;

    if (cfg->outputtype == MCXB_OT_FLUX || cfg->outputtype == MCXB_OT_FLUENCE || adjoint || (cfg->omega > 0.f && !replay)) {
        scale = (float)(cfg->unitinmm / (energytot * Vvox * cfg->tstep));

        if (cfg->outputtype == MCXB_OT_FLUENCE || adjoint) {
            scale *= cfg->tstep;
        }
    } else if (cfg->outputtype == MCXB_OT_ENERGY || cfg->outputtype == MCXB_OT_L) {
        scale = (float)(1.0 / energytot);
    } else if (cfg->replay_weight) {
        /* sensitivity outputs of a replay: unitinmm / sum of the detected weights (src/mcx_host.cpp:1424-1432) */
        scale = 0.f;

        for (uint64_t i = 0; i < cfg->nphoton; i++) {
            scale += cfg->replay_weight[i];
        }

        if (scale > 0.f) {
            scale = cfg->unitinmm / scale;
        }
    }

    if (cfg->extrasrclen && cfg->srcid < 0) {
        scale *= (float)(cfg->extrasrclen + 1);
    }

    return scale;
}

extern "C" float mcxb_normalizer(const mcxb_config* cfg, double energytot) {
    return normalizer_impl(cfg, energytot, cfg->replay_seed != nullptr);
}

/* ---- small device kernels -------------------------------------------------------------------- */
template <typename AccT>
__global__ void finalize_kernel(const AccT* __restrict__ acc, float* __restrict__ out, size_t n, uint32_t copies) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        AccT sum = acc[i];

        for (uint32_t c = 1; c < copies; c++) {
            sum += acc[c * n + i];
        }

        out[i] = (float)sum;
    }
}

/* ---- adjoint post-kernels (mcx_adjoint_kernel / mcx_adjoint_dcoeff_kernel, src/mcx_core.cl:3313-3512).  The reference
 *      re-sums the time gates of a volume for every voxel, every pair and every finite-difference neighbour; here the
 *      continuous-wave volumes are formed once and the pair loop reads them ---- */
__global__ void adjoint_cw_kernel(const float* __restrict__ field, float* __restrict__ cw, uint32_t dimxyz, uint32_t maxgate, uint32_t nslot) {
    const size_t n = (size_t)dimxyz * nslot;

    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t slot = i / dimxyz, vox = i - slot * dimxyz;
        float sum = 0.f;

        for (uint32_t t = 0; t < maxgate; t++) {       /* same order as mcx_cw_sum (:3322-3336) */
            sum += field[vox + (size_t)(t + slot * maxgate) * dimxyz];
        }

        cw[i] = sum;
    }
}

/* second-order finite difference along one axis in voxel units (mcx_fd_grad, :3345-3381) */
__device__ __forceinline__ float adjoint_fd(const float* __restrict__ f, size_t at, uint32_t i, uint32_t n, size_t stride) {
    if (n <= 1u) {
        return 0.f;
    }

    const float f0 = f[at];

    if (i == 0u) {
        const float fp1 = f[at + stride];
        return (n == 2u) ? (fp1 - f0) : (-3.f * f0 + 4.f * fp1 - f[at + 2 * stride]) * 0.5f;
    }

    if (i == n - 1u) {
        const float fm1 = f[at - stride];
        return (n == 2u) ? (f0 - fm1) : (f[at - 2 * stride] - 4.f * fm1 + 3.f * f0) * 0.5f;
    }

    return (f[at + stride] - f[at - stride]) * 0.5f;
}

__global__ void adjoint_pair_kernel(const float* __restrict__ re, const float* __restrict__ im, float* __restrict__ out, uint32_t nx, uint32_t ny,
                                    uint32_t nz, uint32_t ns, uint32_t nd, int gradient) {
    const uint32_t dimxyz = nx * ny * nz;
    const size_t adjointlen = (size_t)dimxyz * ns * nd;

    for (uint32_t vox = blockIdx.x * blockDim.x + threadIdx.x; vox < dimxyz; vox += gridDim.x * blockDim.x) {
        const uint32_t ix = vox % nx, iy = (vox / nx) % ny, iz = vox / (nx * ny);

        for (uint32_t sidx = 0; sidx < ns; sidx++) {
            const size_t sa = (size_t)sidx * dimxyz + vox;
            float sr[3], si[3] = { 0.f, 0.f, 0.f };

            if (gradient) {
                sr[0] = adjoint_fd(re, sa, ix, nx, 1);
                sr[1] = adjoint_fd(re, sa, iy, ny, nx);
                sr[2] = adjoint_fd(re, sa, iz, nz, (size_t)nx * ny);

                if (im) {
                    si[0] = adjoint_fd(im, sa, ix, nx, 1);
                    si[1] = adjoint_fd(im, sa, iy, ny, nx);
                    si[2] = adjoint_fd(im, sa, iz, nz, (size_t)nx * ny);
                }
            } else {
                sr[0] = re[sa];
                sr[1] = sr[2] = 0.f;
                si[0] = im ? im[sa] : 0.f;
            }

            for (uint32_t d = 0; d < nd; d++) {
                const size_t da = (size_t)(ns + d) * dimxyz + vox;
                const size_t o = (size_t)(sidx * nd + d) * dimxyz + vox;
                float dr[3], di[3] = { 0.f, 0.f, 0.f };

                if (gradient) {
                    dr[0] = adjoint_fd(re, da, ix, nx, 1);
                    dr[1] = adjoint_fd(re, da, iy, ny, nx);
                    dr[2] = adjoint_fd(re, da, iz, nz, (size_t)nx * ny);

                    if (im) {
                        di[0] = adjoint_fd(im, da, ix, nx, 1);
                        di[1] = adjoint_fd(im, da, iy, ny, nx);
                        di[2] = adjoint_fd(im, da, iz, nz, (size_t)nx * ny);
                    }
                } else {
                    dr[0] = re[da];
                    dr[1] = dr[2] = 0.f;
                    di[0] = im ? im[da] : 0.f;
                }

                /* complex product / dot product (:3436-3449, 3497-3503), terms in the reference's order */
                float val = gradient ? (__fmul_rn(sr[0], dr[0]) + __fmul_rn(sr[1], dr[1]) + __fmul_rn(sr[2], dr[2])) : __fmul_rn(sr[0], dr[0]);

                if (im) {
                    val -= gradient ? (__fmul_rn(si[0], di[0]) + __fmul_rn(si[1], di[1]) + __fmul_rn(si[2], di[2])) : __fmul_rn(si[0], di[0]);
                    out[o + adjointlen] = gradient ? (__fmul_rn(sr[0], di[0]) + __fmul_rn(sr[1], di[1]) + __fmul_rn(sr[2], di[2])
                                                      + __fmul_rn(si[0], dr[0]) + __fmul_rn(si[1], dr[1]) + __fmul_rn(si[2], dr[2]))
                                          : (__fmul_rn(sr[0], di[0]) + __fmul_rn(si[0], dr[0]));
                }

                out[o] = val;
            }
        }
    }
}

/* MCX_DEBUG_RNG mode of the reference kernel (src/mcx_core.cl:2408-2414) */
__global__ void rngdebug_kernel(const uint32_t* __restrict__ seeds, float* __restrict__ out, uint32_t n, uint32_t nthread) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;

    if (tid >= nthread) {
        return;
    }

    Rng rng;
    rng_seed(rng, seeds + 4 * (size_t)tid);

    for (uint32_t i = tid; i < n; i += nthread) {
        out[i] = rng_uniform(rng);
    }
}

/* ------------------------------------------------------------------------------------------------- */
static void sim_free(mcxb_sim* s) {
    if (!s) {
        return;
    }

    cudaSetDevice(s->device);
    cudaDeviceSynchronize();        /* pooled buffers may be handed out again at once: nothing may still be using them */
    void* bufs[] = { s->d_traj, s->d_trajcount, s->d_media, s->d_field, s->d_field32, s->d_tables, s->d_seeds, s->d_det, s->d_detcount, s->d_seedout,
                     s->d_counter, s->d_energy, s->d_pattern, s->d_invcdf, s->d_stats, s->h_field, s->h_small,
                     s->d_rseed, s->d_rweight, s->d_rtof, s->d_rdetid, s->h_progress, s->d_smatrix
                   };

    for (void* b : bufs) {
        pool().put(b);
    }

    if (s->pollstream) {
        cudaStreamDestroy(s->pollstream);
    }

    if (s->ev0) {
        cudaEventDestroy(s->ev0);
    }

    if (s->ev1) {
        cudaEventDestroy(s->ev1);
    }

    delete s;
}

extern "C" void mcxb_sim_destroy(mcxb_sim* sim) {
    sim_free(sim);
}

static int sim_create_impl(const mcxb_config* cfg, int device, mcxb_sim* s) {
    if (!cfg || cfg->abi_version != MCXB_ABI_VERSION) {
        return fail(MCXB_ERR_ARG, "mcxb_config.abi_version mismatch (library %d)", MCXB_ABI_VERSION);
    }

    if (!cfg->vol || !cfg->prop || cfg->dimx == 0 || cfg->dimy == 0 || cfg->dimz == 0 || cfg->medianum == 0) {
        return fail(MCXB_ERR_ARG, "volume and media table are required");
    }

    for (uint32_t i = 0; i < cfg->medianum; i++) {
        /* a NaN / infinite optical property or a refractive index <= 0 turns positions into NaN, and a packet at NaN never
         * leaves the persistent kernel: refused here rather than hanging the device (the reference does not check) */
        const mcxb_f4& m = cfg->prop[i];

        if (!std::isfinite(m.x) || !std::isfinite(m.y) || !std::isfinite(m.z) || !std::isfinite(m.w) || !(m.w > 0.f)) {
            return fail(MCXB_ERR_ARG, "media row %u {mua %g, mus %g, g %g, n %g} must be finite with n > 0", i, m.x, m.y, m.z, m.w);
        }
    }

    if (cfg->dimx > 32767 || cfg->dimy > 32767 || cfg->dimz > 32767 || (uint64_t)cfg->dimx * cfg->dimy * cfg->dimz >= 0x7FFFFFFFull) {
        return fail(MCXB_ERR_ARG, "grid dimensions exceed the 16-bit voxel coordinates of the photon kernel");
    }

    if (cfg->srctype < 0 || cfg->srctype > MCXB_SRC_RING) {
        return fail(MCXB_ERR_ARG, "the specified source type is not supported");
    }

    if (cfg->tstep <= 0.f || cfg->tend <= cfg->tstart) {
        return fail(MCXB_ERR_ARG, "incorrect time gate settings");
    }

    if (cfg->respin < 0) {
        return fail(MCXB_ERR_ARG, "negative respin is not supported");
    }

    if (cfg->respin > 1 && cfg->replay_seed) {
        return fail(MCXB_ERR_ARG, "respin is disabled in the replay mode");       /* src/mcx_utils.c:1633-1636 */
    }

    if ((cfg->debuglevel & (MCXB_DEBUG_MOVE | MCXB_DEBUG_MOVE_ONLY)) && cfg->maxjumpdebug == 0) {
        return fail(MCXB_ERR_ARG, "trajectory capture needs maxjumpdebug > 0");
    }

    if (cfg->srcnum > 1) {
        /* photon sharing: the reference's host only completes it for the 2-D pattern source (its per-pattern sums
         * srcpw[] are computed for MCX_SRC_PATTERN alone, src/mcx_host.cpp:1351-1364, and dereferenced for both) */
        if (cfg->srctype != MCXB_SRC_PATTERN) {
            return fail(MCXB_ERR_ARG, "photon sharing (srcnum>1) needs the 'pattern' source type");
        }

        if (cfg->extrasrclen) {
            return fail(MCXB_ERR_ARG, "simulating multiple sources currently can not be used with photon-sharing");    /* src/mcx_utils.c:1660 */
        }

        if (cfg->issaveref || cfg->replay_seed) {
            return fail(MCXB_ERR_ARG, "photon sharing cannot be combined with issaveref or replay in this build");
        }

        if (cfg->srcpattern_len < (uint64_t)cfg->src.param1.w * (uint64_t)cfg->src.param2.w * cfg->srcnum) {
            return fail(MCXB_ERR_ARG, "srcpattern is smaller than srcparam1.w x srcparam2.w x srcnum");
        }
    }

    if ((cfg->srctype == MCXB_SRC_PATTERN || cfg->srctype == MCXB_SRC_PATTERN3D) && (!cfg->srcpattern || cfg->srcpattern_len == 0)) {
        return fail(MCXB_ERR_ARG, "pattern sources need srcpattern");
    }

    /* the adjoint types run the forward kernel as a fluence run (src/mcx_core.cl:2844; normalised like fluence,
     * src/mcx_host.cpp:1389-1394); the products are formed afterwards by mcxb_adjoint_products */
    const bool adjoint_ot = cfg->outputtype >= MCXB_OT_ADJOINT && cfg->outputtype <= MCXB_OT_ADJOINT_MUA_MUSP;
    const bool forward_ot = cfg->outputtype == MCXB_OT_FLUX || cfg->outputtype == MCXB_OT_FLUENCE || cfg->outputtype == MCXB_OT_ENERGY ||
                            cfg->outputtype == MCXB_OT_L || adjoint_ot;
    const bool rf_ot = cfg->outputtype == MCXB_OT_RF || cfg->outputtype == MCXB_OT_RFMUS;
    const bool replay_ot = cfg->outputtype == MCXB_OT_JACOBIAN || cfg->outputtype == MCXB_OT_WP || cfg->outputtype == MCXB_OT_DCS ||
                           cfg->outputtype == MCXB_OT_WLTOF || cfg->outputtype == MCXB_OT_WPTOF || rf_ot;
    /* complex packet weights: a modulation frequency in a forward run (src/mcx_host.cpp:473) */
    const bool rfforward = cfg->omega > 0.f && !cfg->replay_seed;
    const bool polarized = cfg->polmedianum > 0;
    const bool svmc = cfg->mediaformat == MCXB_MEDIA_2LABEL_SPLIT;

    if (adjoint_ot && cfg->replay_seed) {
        return fail(MCXB_ERR_ARG, "the adjoint output types are forward runs");
    }

    if (rf_ot && !(cfg->omega > 0.f)) {
        return fail(MCXB_ERR_ARG, "the RF replay outputs need a modulation frequency (omega > 0)");
    }

    if (polarized) {
        /* src/mcx_utils.c:1535-1548: one Mueller matrix per non-background medium of a LABEL volume */
        if (!cfg->smatrix || cfg->medianum != cfg->polmedianum + 1 || cfg->mediaformat > 4) {
            return fail(MCXB_ERR_ARG, "polarised runs need label media and one Mueller matrix (smatrix) per medium: medianum = polmedianum + 1");
        }

        if (cfg->dimx == 1 || cfg->dimy == 1 || cfg->dimz == 1) {
            return fail(MCXB_ERR_ARG, "polarised light is simulated in 3-D domains only");
        }
    }

    if ((rfforward || rf_ot || polarized) && cfg->srcnum > 1) {
        return fail(MCXB_ERR_ARG, "photon sharing cannot be combined with RF or polarised runs in this build");
    }

    if (!forward_ot && !replay_ot) {
        return fail(MCXB_ERR_ARG, "output type %d is outside this build's hot path", cfg->outputtype);
    }

    if (replay_ot && !cfg->replay_seed) {
        return fail(MCXB_ERR_ARG, "output type %d needs photon replay (replay_seed)", cfg->outputtype);
    }

    if (cfg->replay_seed) {
        if (replay_ot && (!cfg->replay_weight || !cfg->replay_tof)) {
            return fail(MCXB_ERR_ARG, "replay outputs need replay_weight and replay_tof");
        }

        if (cfg->replaydet == -1 && replay_ot && (!cfg->replay_detid || cfg->detnum == 0)) {
            return fail(MCXB_ERR_ARG, "replaydet=-1 needs replay_detid and detectors");
        }

        if (cfg->nphoton >= 0xFFFFFFFFull) {
            return fail(MCXB_ERR_ARG, "too many photons to replay");
        }

        if (cfg->seed_skip) {
            return fail(MCXB_ERR_ARG, "replay should only work with a single device");       /* src/mcx_host.cpp:723 */
        }
    }

    if (cfg->extrasrclen && !cfg->srcdata) {
        return fail(MCXB_ERR_ARG, "extrasrclen>0 but srcdata is NULL");
    }

    if (cfg->extrasrclen && cfg->srcid > (int32_t)cfg->extrasrclen + 1) {
        return fail(MCXB_ERR_ARG, "srcid exceeds total defined source count");
    }

    if (cfg->issavedet && cfg->detnum && !cfg->detpos) {
        return fail(MCXB_ERR_ARG, "detnum>0 but detpos is NULL");
    }

    /* positions and directions index the volume: a NaN there is an out-of-bounds voxel address on the device.  (dir.w is
     * the focal length and may be NaN / -inf on purpose: isotropic / Lambertian launch, src/mcx_core.cl:2095-2146) */
    for (uint32_t i = 0; i <= cfg->extrasrclen; i++) {
        const mcxb_source& sr = i ? cfg->srcdata[i - 1] : cfg->src;

        if (!std::isfinite(sr.pos.x) || !std::isfinite(sr.pos.y) || !std::isfinite(sr.pos.z) || !std::isfinite(sr.pos.w) ||
                !std::isfinite(sr.dir.x) || !std::isfinite(sr.dir.y) || !std::isfinite(sr.dir.z)) {
            return fail(MCXB_ERR_ARG, "source %u: position, weight and direction must be finite", i);
        }
    }

    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));

    if (device < 0 || device >= ndev) {
        return fail(MCXB_ERR_NODEVICE, "Specified GPU does not exist");
    }

    CU_TRY(cudaSetDevice(device));
    const cudaDeviceProp* propp = nullptr;
    CU_TRY(device_properties(device, &propp));
    const cudaDeviceProp& prop = *propp;

    if (prop.major < 10) {
        return fail(MCXB_ERR_NODEVICE, "this engine is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    }

    s->device = device;
    s->cfg = *cfg;
    s->cfg.vol = nullptr;
    s->cfg.prop = nullptr;
    s->cfg.srcdata = nullptr;
    s->cfg.srcpattern = nullptr;
    s->cfg.detpos = nullptr;
    s->cfg.invcdf = nullptr;
    s->cfg.angleinvcdf = nullptr;
    s->cfg.replay_seed = nullptr;
    s->cfg.replay_weight = nullptr;
    s->cfg.replay_tof = nullptr;
    s->cfg.replay_detid = nullptr;
    s->cfg.smatrix = nullptr;

    const uint64_t dimxyz = (uint64_t)cfg->dimx * cfg->dimy * cfg->dimz;
    const uint32_t maxgate = count_gates(cfg);

    if (maxgate == 0) {
        return fail(MCXB_ERR_ARG, "incorrect time gate settings");
    }

    /* output volumes: one per pattern with photon sharing, one per source with srcid < 0 (src/mcx_host.cpp:679) */
    const uint32_t nsrcvol = (cfg->srcnum > 1) ? cfg->srcnum : ((cfg->srcid < 0) ? cfg->extrasrclen + 1 : 1);
    s->maxgate = maxgate;
    s->nsrcvol = nsrcvol;
    /* src/mcx_host.cpp:684-689: one volume per detector when every detector is replayed at once */
    s->nrepvol = (cfg->replay_seed && cfg->replaydet == -1) ? std::max(1u, cfg->detnum) : 1u;
    s->rngdebug = (cfg->debuglevel & 1u) != 0;
    /* RF outputs are complex: a second volume set (imaginary parts) follows the first (src/mcx_host.cpp:1263-1276; what
     * pmcxcl allocates, src/pmcxcl.cpp:1209-1211) */
    s->rfplanes = ((rfforward || (rf_ot && cfg->replay_seed)) && !s->rngdebug) ? 2u : 1u;
    s->planelen = dimxyz * maxgate * nsrcvol * s->nrepvol;
    s->fieldlen = s->planelen * s->rfplanes;
    /* adjoint runs (and srcid == -2) launch the detectors appended to the source list as disks (src/mcx_core.cl:2154-2183) */
    const bool detsources = (adjoint_ot || cfg->srcid == -2) && cfg->detnum > 0 && cfg->extrasrclen >= cfg->detnum;
    s->ext = (rfforward || rf_ot || polarized || svmc || detsources) && !s->rngdebug;
    const bool savedet = cfg->issavedet != 0 && !s->rngdebug;
    /* the I flag (Stokes vector of a detected photon) exists in polarised runs only (src/mcx_utils.c:1777-1781) */
    const uint32_t flag = savedet ? (cfg->savedetflag & (polarized ? 0xFFu : 0x7Fu)) : 0u;
    const uint32_t nmed = cfg->medianum - 1;
    const uint32_t partialdata = nmed * ((flag >> 1 & 1u) + (flag >> 2 & 1u) + (flag >> 3 & 1u));
    s->reclen = partialdata + (flag & 1u) + 3 * ((flag >> 4 & 1u) + (flag >> 5 & 1u)) + (flag >> 6 & 1u) + 4 * (flag >> 7 & 1u);

    /* ---- continuous media: the words go to the device as they are (src/mcx_host.cpp:739-745) ---- */
    const bool continuous = cfg->mediaformat > 4;

    if (continuous) {
        const uint32_t f = cfg->mediaformat;

        if ((f < MCXB_MEDIA_LABEL_HALF || f > MCXB_MEDIA_AS_SHORT) && f != MCXB_MEDIA_2LABEL_SPLIT) {
            return fail(MCXB_ERR_ARG, "media format %u (two-word or mixed-label media) is outside this build's hot path", f);
        }

        if (svmc) {
            /* split-voxel media: two words per voxel, labels in the two top bytes of the first (src/mcx_utils.c:1688-1712) */
            for (uint64_t i = 0; i < dimxyz; i++) {
                const uint32_t w = cfg->vol[i] & 0x7FFFFFFFu;

                if ((w >> 24) >= cfg->medianum || ((w >> 16) & 0xFFu) >= cfg->medianum) {
                    return fail(MCXB_ERR_ARG, "input media optical properties are less than the labels in the volume");
                }
            }

            if (cfg->isspecular > 0) {
                /* the reference indexes its media table with the whole media word there (src/mcx_core.cl:1428-1433) */
                return fail(MCXB_ERR_ARG, "isspecular is not available with split-voxel media");
            }
        }

        /* src/mcx_utils.c:1760-1766 */
        if ((f == MCXB_MEDIA_AS_F2H || f == MCXB_MEDIA_MUA_FLOAT || f == MCXB_MEDIA_AS_HALF) && cfg->medianum < 2) {
            return fail(MCXB_ERR_ARG, "the 'prop' field must contain at least 2 rows for the requested media format");
        }

        if ((f == MCXB_MEDIA_ASGN_BYTE || f == MCXB_MEDIA_AS_SHORT) && cfg->medianum < 3) {
            return fail(MCXB_ERR_ARG, "the 'prop' field must contain at least 3 rows for the requested media format");
        }

        if (!svmc && cfg->issavedet && (cfg->savedetflag & 0x0Eu) && !(cfg->debuglevel & 1u)) {
            /* the reference indexes its per-medium rows with the media word itself there (src/mcx_core.cl:2515, 2787) */
            return fail(MCXB_ERR_ARG, "per-medium detector records (savedetflag S, P, M) need label media; use savedetflag 'dxvw' or issavedet=0 with continuous media");
        }

        if (cfg->srcnum > 1 || cfg->replay_seed) {
            return fail(MCXB_ERR_ARG, "photon sharing and replay are not available with continuous media in this build");
        }

        s->media32 = true;
        const uint64_t words = svmc ? 2 * dimxyz : dimxyz;
        CU_TRY(dev_alloc(&s->d_media, device, 4 * words));
        CU_TRY(cudaMemcpy(s->d_media, cfg->vol, 4 * words, cudaMemcpyHostToDevice));
    }

    /* ---- media volume: labels packed to 8 (or 16) bits with the detector flag in the top bit ---- */
    uint32_t maxlabel = 0;

    for (uint64_t i = 0; i < dimxyz && !continuous; i++) {
        maxlabel = std::max(maxlabel, cfg->vol[i] & 0x7FFFFFFFu);
    }

    if (maxlabel >= cfg->medianum) {
        return fail(MCXB_ERR_ARG, "input media optical properties are less than the labels in the volume");
    }

    s->media16 = maxlabel > 127;
    /* mean scattering coefficient per voxel edge over the non-zero voxels (mus is already scaled by unitinmm): decides
     * between the kernels with and without the scattering queue below */
    double mus_voxel = 0.0;
    {
        std::vector<uint64_t> hist(cfg->medianum, 0);

        for (uint64_t i = 0; i < dimxyz && !continuous; i++) {
            hist[cfg->vol[i] & 0x7FFFFFFFu]++;
        }

        double sum = 0.0, cnt = 0.0;

        for (uint32_t m = 1; m < cfg->medianum; m++) {
            sum += (double)hist[m] * cfg->prop[m].y;
            cnt += (double)hist[m];
        }

        mus_voxel = cnt > 0.0 ? sum / cnt : 0.0;
    }

    if (maxlabel > 32767) {
        return fail(MCXB_ERR_ARG, "more than 32767 media labels are not supported");
    }

    if (!continuous) {
        std::vector<uint8_t> packed(dimxyz * (s->media16 ? 2 : 1));

        if (s->media16) {
            uint16_t* p = reinterpret_cast<uint16_t*>(packed.data());

            for (uint64_t i = 0; i < dimxyz; i++) {
                p[i] = (uint16_t)((cfg->vol[i] & 0x7FFFu) | ((cfg->vol[i] & 0x80000000u) ? 0x8000u : 0u));
            }
        } else {
            for (uint64_t i = 0; i < dimxyz; i++) {
                packed[i] = (uint8_t)((cfg->vol[i] & 0x7Fu) | ((cfg->vol[i] & 0x80000000u) ? 0x80u : 0u));
            }
        }

        CU_TRY(dev_alloc(&s->d_media, device, packed.size()));
        CU_TRY(cudaMemcpy(s->d_media, packed.data(), packed.size(), cudaMemcpyHostToDevice));
    }

    /* ---- tables: media, sources (main first), detectors ---- */
    const uint32_t tablen = cfg->medianum + 4 * (1 + cfg->extrasrclen) + cfg->detnum;
    {
        std::vector<float4> tab(tablen);

        for (uint32_t i = 0; i < cfg->medianum; i++) {
            tab[i] = make_float4(cfg->prop[i].x, cfg->prop[i].y, cfg->prop[i].z, cfg->prop[i].w);

            /* mcx_validatecfg (src/mcx_utils.c:1647-1652): a scattering coefficient of exactly 0 becomes EPS, so that a segment
             * length slen / mus is never 0 / 0.  The reference's front-ends have done this already; a direct caller of the C
             * ABI may not have */
            if (tab[i].y == 0.f && !polarized) {
                tab[i].y = 1e-10f;
            }
        }

        /* With gscatter active the reduced coefficient is mus (1 - g), and the conventional background row {0, 0, 1, 1} has
         * g = 1: a packet that lives on inside a label-0 voxel of the grid (reflection compiled out, or matched indices)
         * would then step by 0 / 0 and never leave -- the reference's kernel does not return from that (found by
         * tools/oracle_fuzz.py).  Row 0's g only ever matters for such a packet; it is read as 0 here. */
        if (cfg->gscatter < 1000000000u && tab[0].z >= 1.f) {
            tab[0].z = 0.f;
        }

        auto put = [&](uint32_t at, const mcxb_source & src) {
            tab[at + 0] = make_float4(src.pos.x, src.pos.y, src.pos.z, src.pos.w);
            tab[at + 1] = make_float4(src.dir.x, src.dir.y, src.dir.z, src.dir.w);
            tab[at + 2] = make_float4(src.param1.x, src.param1.y, src.param1.z, src.param1.w);
            tab[at + 3] = make_float4(src.param2.x, src.param2.y, src.param2.z, src.param2.w);
        };
        put(cfg->medianum, cfg->src);

        for (uint32_t i = 0; i < cfg->extrasrclen; i++) {
            put(cfg->medianum + 4 * (i + 1), cfg->srcdata[i]);
        }

        for (uint32_t i = 0; i < cfg->detnum; i++) {
            tab[cfg->medianum + 4 * (1 + cfg->extrasrclen) + i] = make_float4(cfg->detpos[i].x, cfg->detpos[i].y, cfg->detpos[i].z, cfg->detpos[i].w);
        }

        CU_TRY(dev_alloc(&s->d_tables, device, sizeof(float4) * tablen));
        CU_TRY(cudaMemcpy(s->d_tables, tab.data(), sizeof(float4) * tablen, cudaMemcpyHostToDevice));
    }

    if (cfg->srcnum > 1) {
        /* Kahan sums like mcx_kahanSum (src/mcx_host.cpp:1357-1362) */
        const uint64_t psize = (uint64_t)cfg->src.param1.w * (uint64_t)cfg->src.param2.w;
        s->h_srcpw.assign(cfg->srcnum, 0.f);

        for (uint32_t i = 0; i < cfg->srcnum; i++) {
            float sum = 0.f, kc = 0.f;

            for (uint64_t j = 0; j < psize; j++) {
                const float y = cfg->srcpattern[j * cfg->srcnum + i] - kc;
                const float t = sum + y;
                kc = (t - sum) - y;
                sum = t;
            }

            s->h_srcpw[i] = sum;
        }
    }

    if (cfg->srcpattern && cfg->srcpattern_len) {
        CU_TRY(dev_alloc(&s->d_pattern, device, sizeof(float) * cfg->srcpattern_len));
        CU_TRY(cudaMemcpy(s->d_pattern, cfg->srcpattern, sizeof(float) * cfg->srcpattern_len, cudaMemcpyHostToDevice));
    }

    const uint32_t nphase = cfg->invcdf ? cfg->nphase : 0, nangle = cfg->angleinvcdf ? cfg->nangle : 0;

    if (nphase + nangle) {
        std::vector<float> t(nphase + nangle);

        if (nphase) {
            memcpy(t.data(), cfg->invcdf, sizeof(float) * nphase);
        }

        if (nangle) {
            memcpy(t.data() + nphase, cfg->angleinvcdf, sizeof(float) * nangle);
        }

        CU_TRY(dev_alloc(&s->d_invcdf, device, sizeof(float) * t.size()));
        CU_TRY(cudaMemcpy(s->d_invcdf, t.data(), sizeof(float) * t.size(), cudaMemcpyHostToDevice));
    }

    /* ---- Mueller matrices of a polarised run (src/mcx_host.cpp:823-827) ---- */
    if (polarized) {
        const size_t n = (size_t)cfg->polmedianum * MCXB_NANGLES;
        CU_TRY(dev_alloc(&s->d_smatrix, device, sizeof(float4) * n));
        CU_TRY(cudaMemcpy(s->d_smatrix, cfg->smatrix, sizeof(float4) * n, cudaMemcpyHostToDevice));
    }

    /* ---- replay records (src/mcx_host.cpp:722-737) ---- */
    if (cfg->replay_seed) {
        const size_t n = std::max<uint64_t>(cfg->nphoton, 1);
        CU_TRY(dev_alloc(&s->d_rseed, device, 16 * n));
        CU_TRY(cudaMemcpy(s->d_rseed, cfg->replay_seed, 16 * (size_t)cfg->nphoton, cudaMemcpyHostToDevice));

        if (cfg->replay_weight) {
            CU_TRY(dev_alloc(&s->d_rweight, device, 4 * n));
            CU_TRY(cudaMemcpy(s->d_rweight, cfg->replay_weight, 4 * (size_t)cfg->nphoton, cudaMemcpyHostToDevice));
            s->h_rweight.assign(cfg->replay_weight, cfg->replay_weight + cfg->nphoton);
        }

        if (cfg->replay_tof) {
            CU_TRY(dev_alloc(&s->d_rtof, device, 4 * n));
            CU_TRY(cudaMemcpy(s->d_rtof, cfg->replay_tof, 4 * (size_t)cfg->nphoton, cudaMemcpyHostToDevice));
        }

        if (cfg->replay_detid) {
            CU_TRY(dev_alloc(&s->d_rdetid, device, 4 * n));
            CU_TRY(cudaMemcpy(s->d_rdetid, cfg->replay_detid, 4 * (size_t)cfg->nphoton, cudaMemcpyHostToDevice));
            s->h_rdetid.assign(cfg->replay_detid, cfg->replay_detid + cfg->nphoton);
        }
    }

    /* ---- kernel choice and launch shape ---- */
    s->acc64 = cfg->accum != MCXB_ACCUM_F32;
    const bool refl = needs_reflection(cfg);
    const bool stats = (cfg->debuglevel & MCXB_DEBUG_STATS) != 0;
    /* explicit per-face boundary codes / detect-on-face flags: common kernels of their own (BCODES; pencil or any source,
     * 8-bit media, fp64 accumulators, default record), else the generic ones.  MCXB_BC_GENERIC=1: always generic (A/B) */
    bool anybc = false;

    for (int i = 0; i < 12; i++) {
        anybc = anybc || cfg->bc[i] != 0;
    }

    /* the same kernels serve the energy and path-length outputs of an otherwise common configuration */
    const bool richot = cfg->outputtype == MCXB_OT_ENERGY || cfg->outputtype == MCXB_OT_L;
    const bool hasbc = (anybc || richot) && getenv("MCXB_BC_GENERIC") == nullptr;
    const bool common = is_common_config(cfg, savedet, nphase) && !s->media32 && !s->ext && (!(anybc || richot) || hasbc);

    const int detmode = savedet ? (flag == 0x5u ? 1 : 2) : 0;
    /* Scattering queue (photon_kernel.cuh): pays where a packet crosses several voxels per scattering event.  Measured on
     * B200: +8..10 % at mus = 1 per voxel (cube60 / cube60b), +20 % at mus <= 0.2 (skinvessel), -9..-12 % at mus = 8..40
     * (colin27, digimouse).  Not used when seeds are recorded (a replay must see the reference's draw order) or when the
     * record holds anything but the default fields. */
    bool queue = common && !stats && detmode < 2 && !s->media16 && s->acc64 && cfg->issaveseed <= 0 && mus_voxel <= kQueueMaxMus;

    if (const char* e = getenv("MCXB_SCATTER_QUEUE")) {       /* tuning override: 0 = never, 1 = whenever a queue kernel exists */
        queue = queue ? atoi(e) != 0 : (atoi(e) != 0 && common && !stats && detmode < 2 && !s->media16 && s->acc64 && cfg->issaveseed <= 0);
    }

    const KernelEntry* ke = nullptr;
    const int mediabits = s->media32 ? 32 : (s->media16 ? 16 : 8);

    if (stats && s->ext) {
        return fail(MCXB_ERR_ARG, "the instrumented (stats) kernel does not carry the polarised / RF code");
    }

    if (stats && s->media32) {
        return fail(MCXB_ERR_ARG, "the instrumented (stats) kernel exists for label media only");
    }

    if (stats) {
        ke = find_kernel(srcAny, true, 1, mediabits, true, true, false);
    } else {
        if (queue) {
            ke = find_kernel(cfg->srctype, refl, detmode, mediabits, s->acc64, false, common, true, false, hasbc);
        }

        if (!ke) {
            ke = find_kernel(cfg->srctype, refl, detmode, mediabits, s->acc64, false, common, false, s->ext, hasbc);
        }

        if (!ke && s->ext && !s->acc64) {
            s->acc64 = true;        /* the extended-physics kernels exist with fp64 accumulators only */
            ke = find_kernel(cfg->srctype, refl, detmode, mediabits, true, false, common, false, true);
        }
    }

    if (!ke) {
        return fail(MCXB_ERR_ARG, "no kernel specialisation for this configuration");
    }

    if (stats) {
        s->acc64 = true;
    }

    const uint32_t ftablen = (nphase + nangle + 1u) & ~1u;
    int perSM = 0;

    for (int attempt = 0; attempt < 2; attempt++) {
        s->fn = ke->fn;
        s->kname = ke->name;
        s->smem = sizeof(float4) * ke->queue * kBlock      /* scattering queue (photon_kernel.cuh) */
                  + sizeof(float4) * tablen + sizeof(float) * ftablen + (ke->savedet ? sizeof(float) * partialdata * kBlock : 0)
                  + ((savedet && cfg->issaveseed) ? 2 * sizeof(unsigned long long) * kBlock : 0)
                  + (cfg->extrasrclen ? sizeof(int) * kBlock : 0);      /* source id per thread (common kernels) */
        const bool fits = s->smem <= (size_t)prop.sharedMemPerBlockOptin;

        if (fits) {
            if (s->smem > 48 * 1024) {
                CU_TRY(cudaFuncSetAttribute((const void*)s->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)s->smem));
            }

            CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, (const void*)s->fn, kBlock, s->smem));
        }

        if (ke->queue && (!fits || perSM < MCXB_MINBLOCKS)) {
            /* the queue's shared memory would cost resident blocks (many partial-path rows): scatter in place instead */
            ke = find_kernel(cfg->srctype, refl, detmode, mediabits, s->acc64, false, common, false, false, hasbc);

            if (!ke) {
                return fail(MCXB_ERR_ARG, "no kernel specialisation for this configuration");
            }

            continue;
        }

        if (!fits) {
            return fail(MCXB_ERR_NOMEM, "configuration needs %zu bytes of shared memory per block (limit %zu)", s->smem, (size_t)prop.sharedMemPerBlockOptin);
        }

        break;
    }

    if (perSM < 1) {
        return fail(MCXB_ERR_NOMEM, "photon kernel does not fit on an SM");
    }

    if (cfg->nthread) {
        s->nblock = std::max(1u, cfg->nthread / kBlock);
    } else {
        s->nblock = (uint32_t)(prop.multiProcessorCount * perSM);      /* persistent: exactly one resident wave */
    }

    s->nthread = s->nblock * kBlock;

    /* ---- seeds ---- */
    {
        std::vector<uint32_t> seeds;
        cached_seeds(cfg->seed, cfg->seed_skip, s->nthread, seeds);
        CU_TRY(dev_alloc(&s->d_seeds, device, seeds.size() * 4));
        CU_TRY(cudaMemcpy(s->d_seeds, seeds.data(), seeds.size() * 4, cudaMemcpyHostToDevice));
    }

    /* ---- outputs ---- */
    {
        /* accumulator copies: as many as fit a 40 MB budget (a third of the 126 MB L2), at most 8 */
        const uint64_t one = (uint64_t)(s->acc64 ? 8 : 4) * s->fieldlen;
        uint64_t k = std::min<uint64_t>(8, std::max<uint64_t>(1, ((uint64_t)40 << 20) / std::max<uint64_t>(one, 1)));

        if (const char* e = getenv("MCXB_ACC_COPIES")) {
            /* tuning override, bounded: at most 16 copies and at most 1 GB of accumulators */
            k = (uint64_t)std::min(16, std::max(1, atoi(e)));
            k = std::max<uint64_t>(1, std::min<uint64_t>(k, ((uint64_t)1 << 30) / std::max<uint64_t>(one, 1)));
        }

        if (k * s->fieldlen > 0xFFFFFFFFull) {
            k = 1;          /* the kernel keeps the copy offset in 32 bits */
        }

        s->acccopies = s->rngdebug ? 1u : (uint32_t)std::min<uint64_t>(k, s->nblock);
        CU_TRY(dev_alloc(&s->d_field, device, one * s->acccopies));
    }
    CU_TRY(dev_alloc(&s->d_field32, device, 4 * s->fieldlen));
    CU_TRY(dev_alloc(&s->d_energy, device, 2 * sizeof(double)));
    CU_TRY(dev_alloc(&s->d_counter, device, sizeof(unsigned long long)));
    CU_TRY(dev_alloc(&s->d_detcount, device, sizeof(uint32_t)));
    CU_TRY(dev_alloc(&s->d_stats, device, 3 * sizeof(unsigned long long)));

    if ((cfg->debuglevel & (MCXB_DEBUG_MOVE | MCXB_DEBUG_MOVE_ONLY)) && !s->rngdebug) {
        CU_TRY(dev_alloc(&s->d_traj, device, sizeof(float) * MCXB_TRAJ_RECLEN * (size_t)cfg->maxjumpdebug));
        CU_TRY(dev_alloc(&s->d_trajcount, device, sizeof(uint32_t)));
    }

    if (savedet) {
        CU_TRY(dev_alloc(&s->d_det, device, sizeof(float) * std::max<size_t>(1, (size_t)cfg->maxdetphoton * std::max(1u, s->reclen))));

        if (cfg->issaveseed) {
            CU_TRY(dev_alloc(&s->d_seedout, device, 2 * sizeof(unsigned long long) * std::max<size_t>(1, cfg->maxdetphoton)));
        }
    }

    CU_TRY(host_alloc(&s->h_field, 4 * s->fieldlen));
    CU_TRY(host_alloc(&s->h_small, 8 * sizeof(double)));
    CU_TRY(cudaEventCreate(&s->ev0));
    CU_TRY(cudaEventCreate(&s->ev1));

    /* ---- kernel parameters ---- */
    SimParam& P = s->P;
    memset(&P, 0, sizeof(P));
    P.nx = cfg->dimx;
    P.ny = cfg->dimy;
    P.nz = cfg->dimz;
    P.dimxy = cfg->dimx * cfg->dimy;
    P.dimxyz = (uint32_t)dimxyz;
    P.fieldlen = (uint32_t)std::min<uint64_t>(s->fieldlen, 0xFFFFFFFFu);
    P.fnx = (float)cfg->dimx;
    P.fny = (float)cfg->dimy;
    P.fnz = (float)cfg->dimz;
    P.twin0 = cfg->tstart;
    P.twin1 = cfg->tstart + cfg->tstep * maxgate;        /* src/mcx_host.cpp:1076 */
    P.Rtstep = 1.f / cfg->tstep;
    const float R_C0 = 3.335640951981520e-12f;
    P.oneoverc0 = R_C0 * cfg->unitinmm;
    P.minaccumtime = cfg->unitinmm * R_C0 * cfg->unitinmm;
    P.maxgate = maxgate;
    P.minenergy = cfg->minenergy;
    P.gscatter = cfg->gscatter;
    P.doreflect = cfg->isreflect != 0;
    P.save2pt = cfg->issave2pt != 0;
    P.outputtype = adjoint_ot ? (uint32_t)MCXB_OT_FLUENCE : (uint32_t)cfg->outputtype;
    {
        uint32_t is2d = (cfg->dimx == 1 ? 1 : (cfg->dimy == 1 ? 2 : (cfg->dimz == 1 ? 3 : 0)));

        if (is2d) {
            is2d = is2d * (((cfg->dimx > 1) + (cfg->dimy > 1) + (cfg->dimz > 1)) == 2);
        }

        P.is2d = is2d;
    }
    P.isspecular = cfg->isspecular > 0;
    P.issaveref = (uint32_t)cfg->issaveref;
    P.issaveseed = (savedet && cfg->issaveseed > 0) ? 1u : 0u;
    P.voidtime = cfg->voidtime;
    P.maxvoidstep = (uint32_t)cfg->maxvoidstep;
    memcpy(P.bc, cfg->bc, 12);
    P.srctype = cfg->srctype;
    P.srcid = cfg->srcid;
    P.extrasrclen = cfg->extrasrclen;
    P.srcnum = std::max(1u, cfg->srcnum);
    P.medianum = cfg->medianum;
    P.detnum = cfg->detnum;
    P.tablen = tablen;
    P.tables = s->d_tables;
    P.mediaformat = s->media32 ? cfg->mediaformat : 0u;
    {
        /* the launch of the usual pencil beam is a copy of the source record (photon_kernel.cuh, next_packet) */
        uint32_t labelbits = 0;
        memcpy(&labelbits, &cfg->src.param2.w, 4);
        P.plainlaunch = (cfg->srctype == MCXB_SRC_PENCIL && cfg->extrasrclen == 0 && nangle == 0 && cfg->src.dir.w == 0.f &&
                         (labelbits & 0x7FFFFFFFu) != 0u && fabsf(cfg->src.pos.w) > cfg->minenergy && !cfg->replay_seed) ? 1u : 0u;

        if (getenv("MCXB_NO_PLAINLAUNCH")) {       /* tuning / A-B only */
            P.plainlaunch = 0;
        }
    }
    P.srcpattern = s->d_pattern;
    P.nphase = nphase;
    P.nangle = nangle;
    P.ftablen = ftablen;
    P.invcdf = s->d_invcdf;
    P.angleinvcdf = s->d_invcdf ? s->d_invcdf + nphase : nullptr;
    P.savedetflag = flag;
    P.partialdata = partialdata;
    P.reclen = s->reclen;
    P.maxdetphoton = savedet ? cfg->maxdetphoton : 0;
    P.detphoton = s->d_det;
    P.detcount = s->d_detcount;
    P.seedout = s->d_seedout;
    P.nphoton = cfg->nphoton;
    P.counter = s->d_counter;
    P.sched = cfg->sched == MCXB_SCHED_STATIC ? 1 : 0;
    P.threadphoton = (uint32_t)(cfg->nphoton / s->nthread);
    P.oddphoton = (int32_t)(cfg->nphoton - (uint64_t)P.threadphoton * s->nthread);
    {
        /* photons claimed per refill: small enough that the last claims end together, large enough to keep
         * the shared counter cold (about 16 refills per thread) */
        const uint64_t per = cfg->nphoton / ((uint64_t)s->nthread * 16u);
        P.chunk = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(per, 1), 32);
    }
    P.media = s->d_media;
    P.field = s->d_field;
    P.seeds = s->d_seeds;
    P.energy = s->d_energy;
    P.stats = stats ? s->d_stats : nullptr;
    P.replayseed = s->d_rseed;
    P.replayweight = s->d_rweight;
    P.replaytof = s->d_rtof;
    P.replaydetid = s->d_rdetid;
    P.replaydet = cfg->replaydet;
    P.nrepvol = s->nrepvol;
    P.widedep = (maxgate > 1 ? 1u : 0u) | ((cfg->extrasrclen && cfg->srcid < 0) ? 2u : 0u);
    P.acccopies = s->acccopies;
    P.accstride = s->fieldlen;
    P.trajdata = s->d_traj;
    P.trajcount = s->d_trajcount;
    P.maxjumpdebug = cfg->maxjumpdebug;
    P.idbase = 0;
    P.smatrix = s->d_smatrix;
    P.maxpolmedia = polarized ? cfg->polmedianum : 0u;
    P.s0i = cfg->srciquv.x;
    P.s0q = cfg->srciquv.y;
    P.s0u = cfg->srciquv.z;
    P.s0v = cfg->srciquv.w;
    P.omega = cfg->omega;
    P.rfforward = rfforward ? 1u : 0u;
    P.rfplane = (s->rfplanes == 2) ? s->planelen : 0ull;
    P.adjfirstdet = detsources ? (cfg->extrasrclen + 1 - cfg->detnum) : 0xFFFFFFFFu;

    if (getenv("MCXB_DEBUG_PTRS")) {
        fprintf(stderr, "mcxb buffers: media %p field %p field32 %p tables %p seeds %p det %p detcount %p counter %p energy %p\n",
                s->d_media, s->d_field, (void*)s->d_field32, (void*)s->d_tables, (void*)s->d_seeds, (void*)s->d_det, (void*)s->d_detcount,
                (void*)s->d_counter, (void*)s->d_energy);
    }

    return MCXB_OK;
}

extern "C" int mcxb_sim_create(const mcxb_config* cfg, int device, mcxb_sim** sim) {
    if (!sim) {
        return fail(MCXB_ERR_ARG, "sim is NULL");
    }

    *sim = nullptr;
    mcxb_sim* s = new mcxb_sim();
    const int rc = sim_create_impl(cfg, device, s);

    if (rc != MCXB_OK) {
        sim_free(s);
        return rc;
    }

    *sim = s;
    return MCXB_OK;
}

extern "C" int mcxb_sim_reset(mcxb_sim* s, void* cuda_stream) {
    if (!s) {
        return fail(MCXB_ERR_ARG, "sim is NULL");
    }

    cudaStream_t st = (cudaStream_t)cuda_stream;
    CU_TRY(cudaSetDevice(s->device));
    CU_TRY(cudaMemsetAsync(s->d_field, 0, (s->acc64 ? 8 : 4) * s->fieldlen * s->acccopies, st));
    CU_TRY(cudaMemsetAsync(s->d_energy, 0, 2 * sizeof(double), st));
    CU_TRY(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned long long), st));
    CU_TRY(cudaMemsetAsync(s->d_detcount, 0, sizeof(uint32_t), st));
    CU_TRY(cudaMemsetAsync(s->d_stats, 0, 3 * sizeof(unsigned long long), st));

    if (s->d_trajcount) {
        CU_TRY(cudaMemsetAsync(s->d_trajcount, 0, sizeof(uint32_t), st));
    }

    s->finalized = false;
    s->launched = false;
    s->launched_photons = 0;
    return MCXB_OK;
}

extern "C" int mcxb_sim_set_photons(mcxb_sim* s, uint64_t nphoton) {
    if (!s) {
        return fail(MCXB_ERR_ARG, "sim is NULL");
    }

    if (s->d_rseed && nphoton > s->cfg.nphoton) {
        return fail(MCXB_ERR_ARG, "a replay cannot run more photons than it has records");
    }

    s->cfg.nphoton = nphoton;
    s->P.nphoton = nphoton;
    s->P.threadphoton = (uint32_t)(nphoton / s->nthread);
    s->P.oddphoton = (int32_t)(nphoton - (uint64_t)s->P.threadphoton * s->nthread);
    const uint64_t per = nphoton / ((uint64_t)s->nthread * 16u);
    s->P.chunk = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(per, 1), 32);
    return MCXB_OK;
}

extern "C" int mcxb_sim_reseed(mcxb_sim* s, int32_t seed, uint64_t seed_skip) {
    if (!s) {
        return fail(MCXB_ERR_ARG, "sim is NULL");
    }

    CU_TRY(cudaSetDevice(s->device));
    std::vector<uint32_t> seeds;
    cached_seeds(seed, seed_skip, s->nthread, seeds);
    CU_TRY(cudaMemcpy(s->d_seeds, seeds.data(), seeds.size() * 4, cudaMemcpyHostToDevice));
    s->cfg.seed = seed;
    s->cfg.seed_skip = seed_skip;
    return MCXB_OK;
}

extern "C" int mcxb_sim_launch(mcxb_sim* s, void* cuda_stream) {
    if (!s) {
        return fail(MCXB_ERR_ARG, "sim is NULL");
    }

    cudaStream_t st = (cudaStream_t)cuda_stream;
    CU_TRY(cudaSetDevice(s->device));

    if (s->launched) {
        /* another batch on top of the accumulated results (respin): the photon counter starts over, everything else
         * (volume, energies, detected photons, trajectories) keeps accumulating */
        CU_TRY(cudaMemsetAsync(s->d_counter, 0, sizeof(unsigned long long), st));
        s->finalized = false;
    }

    s->P.idbase = (uint32_t)s->launched_photons;
    CU_TRY(cudaEventRecord(s->ev0, st));

    if (s->rngdebug) {
        rngdebug_kernel <<< s->nblock, kBlock, 0, st>>>(s->d_seeds, s->d_field32, (uint32_t)s->fieldlen, s->nthread);
        s->finalized = true;
    } else {
        void* args[] = { (void*)& s->P };
        CU_TRY(cudaLaunchKernel((const void*)s->fn, dim3(s->nblock), dim3(kBlock), args, s->smem, st));
    }

    CU_TRY(cudaGetLastError());
    CU_TRY(cudaEventRecord(s->ev1, st));
    s->launches++;
    s->launched = true;
    s->launched_photons += s->P.nphoton;
    return MCXB_OK;
}

extern "C" int mcxb_sim_progress(mcxb_sim* s, uint64_t* claimed, int* finished) {
    if (!s) {
        return fail(MCXB_ERR_ARG, "sim is NULL");
    }

    CU_TRY(cudaSetDevice(s->device));

    if (!s->pollstream) {
        /* a non-blocking stream: its copies overlap the running kernel instead of queueing behind it */
        CU_TRY(cudaStreamCreateWithFlags(&s->pollstream, cudaStreamNonBlocking));
        CU_TRY(host_alloc(&s->h_progress, sizeof(unsigned long long)));
    }

    if (finished) {
        *finished = 1;

        if (s->launched) {
            const cudaError_t e = cudaEventQuery(s->ev1);

            if (e == cudaErrorNotReady) {
                *finished = 0;
            } else if (e != cudaSuccess) {
                return fail(MCXB_ERR_CUDA_BASE - (int)e, "cudaEventQuery failed: %s", cudaGetErrorString(e));
            }
        }
    }

    if (claimed) {
        CU_TRY(cudaMemcpyAsync(s->h_progress, s->d_counter, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->pollstream));
        CU_TRY(cudaStreamSynchronize(s->pollstream));
        /* the counter runs past nphoton by one refill per thread at the end; static scheduling does not use it */
        *claimed = (s->P.sched == 1) ? ((finished && *finished) ? s->P.nphoton : 0) : std::min<uint64_t>(*s->h_progress, s->P.nphoton);
    }

    return MCXB_OK;
}

extern "C" int mcxb_sim_finalize(mcxb_sim* s, void* cuda_stream) {
    if (!s) {
        return fail(MCXB_ERR_ARG, "sim is NULL");
    }

    if (s->finalized) {
        return MCXB_OK;
    }

    cudaStream_t st = (cudaStream_t)cuda_stream;
    CU_TRY(cudaSetDevice(s->device));
    const int grid = (int)std::min<uint64_t>((s->fieldlen + 255) / 256, 148 * 8);

    if (s->acc64) {
        finalize_kernel<double> <<< grid, 256, 0, st>>>((const double*)s->d_field, s->d_field32, s->fieldlen, s->acccopies);
    } else {
        finalize_kernel<float> <<< grid, 256, 0, st>>>((const float*)s->d_field, s->d_field32, s->fieldlen, s->acccopies);
    }

    CU_TRY(cudaGetLastError());
    s->launches++;
    s->finalized = true;
    return MCXB_OK;
}

/* host loops over the whole output (accumulate + normalise, src/mcx_host.cpp:1292-1296 and mcx_normalize): split over a few
 * threads once the volume is large (time-gated atlases reach hundreds of millions of elements, where one core needs ~0.5 s) */
template <typename F> static void for_range(uint64_t n, F body) {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned nt = (n < (4u << 20)) ? 1u : std::min(8u, hw);

    if (nt == 1) {
        body((uint64_t)0, n);
        return;
    }

    std::vector<std::thread> th;
    const uint64_t chunk = (n + nt - 1) / nt;

    for (unsigned t = 0; t < nt; t++) {
        const uint64_t a = t * chunk, b = std::min(n, a + chunk);

        if (a < b) {
            th.emplace_back([ =, &body] { body(a, b); });
        }
    }

    for (auto& t : th) {
        t.join();
    }
}

extern "C" int mcxb_sim_fetch(mcxb_sim* s, void* cuda_stream, mcxb_output* out) {
    if (!s || !out) {
        return fail(MCXB_ERR_ARG, "sim/out is NULL");
    }

    cudaStream_t st = (cudaStream_t)cuda_stream;
    CU_TRY(cudaSetDevice(s->device));
    int rc = mcxb_sim_finalize(s, st);

    if (rc != MCXB_OK) {
        return rc;
    }

    const bool wantfield = out->field != nullptr && s->cfg.issave2pt;

    if (out->field && out->fieldlen < s->fieldlen) {
        return fail(MCXB_ERR_ARG, "field buffer too small: %llu < %llu", (unsigned long long)out->fieldlen, (unsigned long long)s->fieldlen);
    }

    if (wantfield) {
        CU_TRY(cudaMemcpyAsync(s->h_field, s->d_field32, 4 * s->fieldlen, cudaMemcpyDeviceToHost, st));
    }

    CU_TRY(cudaMemcpyAsync(s->h_small, s->d_energy, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(s->h_small + 2, s->d_detcount, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(s->h_small + 3, s->d_stats, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));

    if (s->launched) {
        CU_TRY(cudaEventElapsedTime(&s->last_ms, s->ev0, s->ev1));
    }

    uint32_t detected = 0;
    memcpy(&detected, s->h_small + 2, sizeof(uint32_t));
    out->energyesc = s->h_small[0];
    out->energytot = s->h_small[1];
    out->energyabs = out->energytot - out->energyesc;
    out->detected = detected;
    out->saved = s->d_det ? std::min(detected, s->cfg.maxdetphoton) : 0;
    out->reclen = s->reclen;
    out->maxgate = s->maxgate;
    out->fieldlen = s->fieldlen;
    out->runtime_ms = s->last_ms;
    out->nthread = s->nthread;
    out->nblocksize = kBlock;
    out->kernel_launches = s->launches;
    memcpy(out->stats, s->h_small + 3, 3 * sizeof(unsigned long long));

    if (out->saved && out->detphoton) {
        CU_TRY(cudaMemcpy(out->detphoton, s->d_det, sizeof(float) * (size_t)out->saved * s->reclen, cudaMemcpyDeviceToHost));
    }

    if (out->saved && out->seeddata && s->d_seedout) {
        CU_TRY(cudaMemcpy(out->seeddata, s->d_seedout, 2 * sizeof(unsigned long long) * (size_t)out->saved, cudaMemcpyDeviceToHost));
    }

    out->debugrecorded = out->debugdatalen = 0;

    if (s->d_trajcount) {
        uint32_t n = 0;
        CU_TRY(cudaMemcpy(&n, s->d_trajcount, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        out->debugrecorded = n;
        out->debugdatalen = std::min(n, s->cfg.maxjumpdebug);

        if (out->debugdata && out->debugdatalen) {
            CU_TRY(cudaMemcpy(out->debugdata, s->d_traj, sizeof(float) * MCXB_TRAJ_RECLEN * (size_t)out->debugdatalen, cudaMemcpyDeviceToHost));
        }
    }

    out->normalizer = 1.f;

    if (wantfield) {
        /* cfg->exportfield[i] += field[i]  (src/mcx_host.cpp:1292-1296), then mcx_normalize (src/mcx_utils.c:1208-1218) */
        float* dst = out->field;
        const float* src = s->h_field;
        const uint64_t n = s->fieldlen;

        const bool sens = s->d_rseed && s->cfg.outputtype > MCXB_OT_ENERGY && s->cfg.outputtype != MCXB_OT_L;

        if (s->cfg.isnormalized && sens && s->cfg.replaydet == -1) {
            /* every detector replayed at once: each detector's volumes are scaled by unitinmm / (sum of the weights
             * detected THERE) (src/mcx_host.cpp:1398-1421) */
            const uint64_t block = (uint64_t)s->P.dimxyz * s->maxgate;

            for (uint64_t v = 0; v < (uint64_t)s->nsrcvol * s->nrepvol * s->rfplanes; v++) {
                const int det = (int)(v % s->nrepvol) + 1;
                float scale = 0.f;

                for (size_t i = 0; i < s->h_rweight.size() && i < s->P.nphoton; i++) {
                    if ((s->h_rdetid[i] & 0xFFFF) == det) {
                        scale += s->h_rweight[i];
                    }
                }

                if (scale > 0.f) {
                    scale = s->cfg.unitinmm / scale;
                }

                out->normalizer = scale;

                for (uint64_t i = v * block; i < (v + 1) * block; i++) {
                    dst[i] = (dst[i] + src[i]) * scale;
                }
            }
        } else if (s->cfg.isnormalized && s->cfg.srcnum > 1 && out->energytot > 0.0) {
            /* photon sharing: pattern i is scaled by psize / sum(pattern i) (src/mcx_host.cpp:1436-1447); the volumes
             * are interleaved pattern-fastest, as mcx_normalize(field, scale, n, op, i, srcnum) walks them */
            const float ref = normalizer_impl(&s->cfg, out->energytot, s->d_rseed != nullptr);
            const float psize = (float)((int)s->cfg.src.param1.w * (int)s->cfg.src.param2.w);
            const uint32_t ns = s->cfg.srcnum;
            std::vector<float> scale(ns);

            for (uint32_t i = 0; i < ns; i++) {
                scale[i] = psize / s->h_srcpw[i] * ref;
            }

            out->normalizer = scale[0];

            for (uint64_t i = 0; i < n; i++) {
                dst[i] = (dst[i] + src[i]) * scale[i % ns];
            }
        } else if (s->cfg.isnormalized && !s->rngdebug && out->energytot > 0.0) {     /* nothing launched: leave the zeros alone */
            mcxb_config nc = s->cfg;
            nc.replay_weight = (sens && !s->h_rweight.empty()) ? s->h_rweight.data() : nullptr;
            nc.nphoton = s->P.nphoton;
            const float scale = normalizer_impl(&nc, out->energytot, s->d_rseed != nullptr);
            out->normalizer = scale;

            for_range(n, [&](uint64_t a, uint64_t b) {
                for (uint64_t i = a; i < b; i++) {
                    dst[i] = (dst[i] + src[i]) * scale;
                }
            });
        } else {
            for_range(n, [&](uint64_t a, uint64_t b) {
                for (uint64_t i = a; i < b; i++) {
                    dst[i] += src[i];
                }
            });
        }
    }

    return MCXB_OK;
}

extern "C" int mcxb_adjoint_products(int device, const float* field_re, const float* field_im, uint32_t dimx, uint32_t dimy, uint32_t dimz,
                                     uint32_t maxgate, uint32_t ns, uint32_t nd, int gradient, float* out) {
    if (!field_re || !out || dimx == 0 || dimy == 0 || dimz == 0 || maxgate == 0 || ns == 0 || nd == 0) {
        return fail(MCXB_ERR_ARG, "mcxb_adjoint_products: fluence volumes of at least one source and one detector are required");
    }

    const uint64_t dimxyz = (uint64_t)dimx * dimy * dimz;

    if (dimxyz * (ns + nd) * maxgate >= 0xFFFFFFFFull || dimxyz * ns * nd >= 0xFFFFFFFFull) {
        return fail(MCXB_ERR_ARG, "mcxb_adjoint_products: volumes too large");
    }

    int ndev = 0;
    CU_TRY(cudaGetDeviceCount(&ndev));

    if (device < 0 || device >= ndev) {
        return fail(MCXB_ERR_NODEVICE, "Specified GPU does not exist");
    }

    CU_TRY(cudaSetDevice(device));
    const size_t inlen = dimxyz * maxgate * (ns + nd), cwlen = dimxyz * (ns + nd), pairs = dimxyz * ns * nd;
    const int planes = field_im ? 2 : 1;
    float* d_in = nullptr;
    float* d_cw = nullptr;
    float* d_out = nullptr;
    int rc = MCXB_OK;
    auto release = [&]() {
        pool().put(d_in);
        pool().put(d_cw);
        pool().put(d_out);
    };
#define ADJ_TRY(call)                                                                                             \
    do {                                                                                                          \
        cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess) {                                                                                 \
            rc = fail(MCXB_ERR_CUDA_BASE - (int)e__, "%s failed: %s", #call, cudaGetErrorString(e__));              \
            release();                                                                                            \
            return rc;                                                                                            \
        }                                                                                                         \
    } while (0)
    ADJ_TRY(dev_alloc(&d_in, device, sizeof(float) * inlen));
    ADJ_TRY(dev_alloc(&d_cw, device, sizeof(float) * cwlen * planes));
    ADJ_TRY(dev_alloc(&d_out, device, sizeof(float) * pairs * planes));
    const int grid = (int)std::min<uint64_t>((cwlen + 255) / 256, 148 * 8);

    for (int p = 0; p < planes; p++) {
        ADJ_TRY(cudaMemcpy(d_in, p ? field_im : field_re, sizeof(float) * inlen, cudaMemcpyHostToDevice));
        adjoint_cw_kernel <<< grid, 256>>>(d_in, d_cw + (size_t)p * cwlen, (uint32_t)dimxyz, maxgate, ns + nd);
        ADJ_TRY(cudaGetLastError());
    }

    adjoint_pair_kernel <<< (int)std::min<uint64_t>((dimxyz + 255) / 256, 148 * 8), 256>>>(d_cw, field_im ? d_cw + cwlen : nullptr, d_out, dimx, dimy, dimz, ns, nd, gradient);
    ADJ_TRY(cudaGetLastError());
    ADJ_TRY(cudaMemcpy(out, d_out, sizeof(float) * pairs * planes, cudaMemcpyDeviceToHost));
#undef ADJ_TRY
    release();
    return MCXB_OK;
}

extern "C" void* mcxb_sim_field_devptr(mcxb_sim* s) {
    return s ? s->d_field32 : nullptr;
}
extern "C" void* mcxb_sim_energy_devptr(mcxb_sim* s) {
    return s ? s->d_energy : nullptr;
}
extern "C" void* mcxb_sim_detphoton_devptr(mcxb_sim* s) {
    return s ? s->d_det : nullptr;
}
extern "C" void* mcxb_sim_detcount_devptr(mcxb_sim* s) {
    return s ? s->d_detcount : nullptr;
}
extern "C" void* mcxb_sim_seeddata_devptr(mcxb_sim* s) {
    return s ? s->d_seedout : nullptr;
}
extern "C" uint64_t mcxb_sim_fieldlen(mcxb_sim* s) {
    return s ? s->fieldlen : 0;
}
extern "C" uint32_t mcxb_sim_reclen(mcxb_sim* s) {
    return s ? s->reclen : 0;
}
extern "C" uint32_t mcxb_sim_acc_copies(mcxb_sim* s) {
    return s ? s->acccopies : 0;
}
extern "C" uint32_t mcxb_sim_nthread(mcxb_sim* s) {
    return s ? s->nthread : 0;
}
extern "C" const char* mcxb_sim_kernel_name(mcxb_sim* s) {
    return s ? s->kname : "";
}
extern "C" float mcxb_sim_last_kernel_ms(mcxb_sim* s) {
    if (!s || !s->launched) {
        return 0.f;
    }

    float ms = 0.f;

    if (cudaEventSynchronize(s->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, s->ev0, s->ev1) == cudaSuccess) {
        s->last_ms = ms;
    }

    return s->last_ms;
}

extern "C" int mcxb_sim_run_batches(mcxb_sim* s, uint64_t nphoton, uint32_t respin, int32_t seed, uint64_t seed_skip, uint64_t seed_stride, float* kernel_ms) {
    /* nphoton split into `respin` batches (remainder to the first ones), batch k seeded from record
     * seed_skip + k * seed_stride of the seed stream: with seed_stride = the threads of ALL devices of the job every
     * batch of every device gets its own slice, in the order the reference draws them (src/mcx_host.cpp:1319-1332) */
    if (!s || respin == 0) {
        return fail(MCXB_ERR_ARG, "sim is NULL or respin is 0");
    }

    float total = 0.f;

    for (uint32_t k = 0; k < respin; k++) {
        const uint64_t batch = nphoton / respin + (k < nphoton % respin ? 1 : 0);
        int rc = mcxb_sim_set_photons(s, batch);

        if (rc == MCXB_OK && k > 0) {
            rc = mcxb_sim_reseed(s, seed, seed_skip + (uint64_t)k * seed_stride);
        }

        if (rc == MCXB_OK) {
            rc = mcxb_sim_launch(s, nullptr);
        }

        if (rc != MCXB_OK) {
            return rc;
        }

        total += mcxb_sim_last_kernel_ms(s);        /* waits for this batch */
    }

    s->cfg.nphoton = nphoton;
    s->P.nphoton = nphoton;

    if (kernel_ms) {
        *kernel_ms = total;
    }

    return MCXB_OK;
}

extern "C" int mcxb_run_simulation(const mcxb_config* cfg, int device, mcxb_output* out) {
    /* MCXB_TIMING=1 prints where the host time of one call goes (create / reset+launch / fetch / destroy) */
    static const bool timing = getenv("MCXB_TIMING") != nullptr;
    typedef std::chrono::steady_clock clk;
    clk::time_point t[5];
    mcxb_sim* s = nullptr;
    t[0] = clk::now();
    int rc = mcxb_sim_create(cfg, device, &s);
    t[1] = clk::now();

    if (rc == MCXB_OK) {
        rc = mcxb_sim_reset(s, nullptr);
    }

    float kernel_ms = 0.f;
    const uint32_t respin = (uint32_t)std::max(1, cfg->respin);

    if (rc == MCXB_OK && respin == 1) {
        rc = mcxb_sim_launch(s, nullptr);
    } else if (rc == MCXB_OK) {
        rc = mcxb_sim_run_batches(s, cfg->nphoton, respin, cfg->seed, cfg->seed_skip, mcxb_sim_nthread(s), &kernel_ms);
    }

    t[2] = clk::now();

    if (rc == MCXB_OK) {
        rc = mcxb_sim_fetch(s, nullptr, out);

        if (respin > 1 && out) {
            out->runtime_ms = kernel_ms;        /* the kernel windows of all batches */
        }
    }

    t[3] = clk::now();
    mcxb_sim_destroy(s);
    t[4] = clk::now();

    if (timing) {
        auto ms = [&](int a, int b) {
            return std::chrono::duration<double, std::milli>(t[b] - t[a]).count();
        };
        fprintf(stderr, "mcxb_run_simulation: create %.2f launch %.2f fetch %.2f (kernel %.2f) destroy %.2f ms\n", ms(0, 1), ms(1, 2), ms(2, 3),
                out ? out->runtime_ms : 0.f, ms(3, 4));
    }

    return rc;
}
