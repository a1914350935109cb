/*
 * engine_multi.cu -- one simulation sharded over several GPUs of one box from ONE host process, behind the C ABI.
 *
 * Replaces the multi-device branch of the reference's mcx_run_simulation: the workload split (src/mcx_host.cpp:650-662,
 * 1011-1012), one slice of the single rand() stream per device (:759-768), the per-device launches inside one timing
 * window (:1098-1168) and -- instead of reading every device back and summing on the host (:1218-1232, 1292-1306) --
 *
 *     ncclReduce   float32 volume          -> device 0      (NVLink / NVSwitch)
 *     ncclReduce   float64 {escaped, launched} -> device 0
 *     ncclAllGather of the detected-photon counts, exclusive scan on the host,
 *     grouped ncclSend / ncclRecv of count_r x reclen floats (and count_r x 16 bytes of RNG states) into the tail of
 *     device 0's record buffer,
 *
 * then ONE device-to-host copy and one normalisation with the global launched energy.  Built on the public staged API
 * (mcxb_sim_*), so a device runs exactly the kernels a single-GPU call runs.
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2: the copy already loaded in the process, e.g. PyTorch's, or the
 * system one), so single-GPU users of libmcxb200.so need no NCCL at all.
 *
 * Peer-memory exchange.  One host process can do better than a collective library for this pattern: when devices[0] can
 * map the memory of every other device (NVLink / NVSwitch: always on a B200 box), ONE kernel on devices[0] reads the
 * float32 volumes of all peers straight over NVLink and adds them into its own (peer_sum_kernel: the gather and the
 * reduction are the same loads), the records are copied peer to peer into the tail of devices[0]'s buffer, and the
 * 16-byte energy pairs and the counts go through the host.  No communicator has to be built -- ncclCommInitAll plus
 * the first collective cost seconds, more than the photon kernels of a 1e9-photon colin27 run on 8 GPUs (measured:
 * profiles/README.md) -- and the sum has a fixed order.  This is the default; MCXB_MULTI_EXCHANGE=nccl selects the NCCL
 * path, which is also the fallback when a peer cannot be mapped.
 */
#include "../../include/mcxb200.h"

#include <cuda_runtime.h>
#include <nccl.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

extern "C" void mcxb_set_last_error(const char* msg);     /* engine.cu */

namespace {

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    mcxb_set_last_error(buf);
    return code;
}

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string why;
};

Nccl& nccl() {
    static Nccl* n = [] {
        Nccl* x = new Nccl();
        const char* names[] = { "libnccl.so.2", "libnccl.so" };

        for (const char* nm : names) {
            x->handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);

            if (x->handle) {
                break;
            }
        }

        if (!x->handle) {
            const char* why = dlerror();        /* a second call would return NULL: dlerror() clears the message it hands out */
            x->why = std::string("libnccl.so.2 not found: ") + (why ? why : "");
            return x;
        }

#define MCXB_SYM(field, name)                                               \
    *(void**)(&x->field) = dlsym(x->handle, name);                          \
    if (!x->field) {                                                        \
        x->why = std::string("symbol missing in libnccl: ") + name;         \
        x->handle = nullptr;                                                \
        return x;                                                           \
    }
        MCXB_SYM(CommInitAll, "ncclCommInitAll")
        MCXB_SYM(GroupStart, "ncclGroupStart")
        MCXB_SYM(GroupEnd, "ncclGroupEnd")
        MCXB_SYM(Reduce, "ncclReduce")
        MCXB_SYM(AllGather, "ncclAllGather")
        MCXB_SYM(Send, "ncclSend")
        MCXB_SYM(Recv, "ncclRecv")
        MCXB_SYM(GetErrorString, "ncclGetErrorString")
        MCXB_SYM(GetVersion, "ncclGetVersion")
#undef MCXB_SYM
        return x;
    }();
    return *n;
}

/* communicators are expensive to build (hundreds of ms): one set per device list per process, kept for its lifetime */
struct CommSet {
    std::vector<ncclComm_t> comm;
    std::vector<cudaStream_t> stream;
    std::vector<uint32_t*> counts;       /* per device: ndev gathered detected-photon counts */
};

int get_comms(const std::vector<int>& devs, CommSet** out) {
    static std::mutex mu;
    static std::map<std::vector<int>, CommSet*> cache;
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(devs);

    if (it != cache.end()) {
        *out = it->second;
        return MCXB_OK;
    }

    Nccl& N = nccl();

    if (!N.handle) {
        return fail(MCXB_ERR_ARG, "multi-GPU runs need NCCL: %s", N.why.c_str());
    }

    CommSet* cs = new CommSet();
    cs->comm.resize(devs.size());
    const ncclResult_t r = N.CommInitAll(cs->comm.data(), (int)devs.size(), devs.data());

    if (r != ncclSuccess) {
        delete cs;
        return fail(MCXB_ERR_CUDA_BASE, "ncclCommInitAll failed: %s", N.GetErrorString(r));
    }

    cs->stream.resize(devs.size());
    cs->counts.resize(devs.size());

    for (size_t i = 0; i < devs.size(); i++) {
        cudaSetDevice(devs[i]);

        if (cudaStreamCreateWithFlags(&cs->stream[i], cudaStreamNonBlocking) != cudaSuccess ||
                cudaMalloc(&cs->counts[i], sizeof(uint32_t) * devs.size()) != cudaSuccess) {
            /* the communicators stay alive (ncclCommDestroy on a half-built set can hang); the set is simply not cached */
            return fail(MCXB_ERR_NOMEM, "cannot allocate the multi-GPU exchange buffers on device %d", devs[i]);
        }
    }

    cache[devs] = cs;
    *out = cs;
    return MCXB_OK;
}

/* out[i] += sum over peers of src[p][i]: the peers' volumes are read over NVLink (peer-mapped pointers), 128 bits at a time */
struct PeerList {
    const float* src[MCXB_MAX_DEVICES];
    int n;
};

__global__ void peer_sum_kernel(float* __restrict__ out, PeerList peers, size_t n) {
    const size_t n4 = n / 4;
    float4* out4 = reinterpret_cast<float4*>(out);

    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = out4[i];

        for (int p = 0; p < peers.n; p++) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(peers.src[p]) + i);
            acc.x += v.x;
            acc.y += v.y;
            acc.z += v.z;
            acc.w += v.w;
        }

        out4[i] = acc;
    }

    for (size_t i = 4 * n4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float acc = out[i];

        for (int p = 0; p < peers.n; p++) {
            acc += peers.src[p][i];
        }

        out[i] = acc;
    }
}

/* can devices[0] map every other device's memory?  (enables the mappings on first use) */
bool enable_peers(const std::vector<int>& devs) {
    if (cudaSetDevice(devs[0]) != cudaSuccess) {
        return false;
    }

    for (size_t i = 1; i < devs.size(); i++) {
        int ok = 0;

        if (cudaDeviceCanAccessPeer(&ok, devs[0], devs[i]) != cudaSuccess || !ok) {
            return false;
        }

        const cudaError_t e = cudaDeviceEnablePeerAccess(devs[i], 0);

        if (e == cudaErrorPeerAccessAlreadyEnabled) {
            cudaGetLastError();
        } else if (e != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
    }

    return true;
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define NCCL_TRY(call)                                                                                            \
    do {                                                                                                          \
        ncclResult_t r__ = (call);                                                                                \
        if (r__ != ncclSuccess) {                                                                                 \
            rc = fail(MCXB_ERR_CUDA_BASE, "%s failed: %s (%s:%d)", #call, N.GetErrorString(r__), __FILE__, __LINE__); \
            goto done;                                                                                            \
        }                                                                                                         \
    } while (0)
#define CUDA_TRY(call)                                                                                            \
    do {                                                                                                          \
        cudaError_t e__ = (call);                                                                                 \
        if (e__ != cudaSuccess) {                                                                                 \
            rc = fail(MCXB_ERR_CUDA_BASE - (int)e__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            goto done;                                                                                            \
        }                                                                                                         \
    } while (0)

}  // namespace

extern "C" int mcxb_nccl_version(void) {
    Nccl& N = nccl();
    int v = 0;

    if (!N.handle || N.GetVersion(&v) != ncclSuccess) {
        return 0;
    }

    return v;
}

extern "C" void mcxb_split_photons(uint64_t nphoton, const float* workload, int ndev, uint64_t* share) {
    /* nphoton * w_i / sum(w), rounded down; the remainder goes to the first devices with a non-zero weight so that
     * exactly nphoton packets are launched (the reference derives threadphoton / oddphoton per device from the same
     * ratio, src/mcx_host.cpp:1011-1012) */
    double full = 0.0;

    for (int i = 0; i < ndev; i++) {
        full += (workload && workload[i] > 0.f) ? workload[i] : 0.0;
    }

    uint64_t assigned = 0;

    for (int i = 0; i < ndev; i++) {
        const double w = (full > 0.0) ? ((workload[i] > 0.f) ? workload[i] : 0.0) : 1.0;
        share[i] = (uint64_t)((double)nphoton * w / (full > 0.0 ? full : (double)ndev));
        assigned += share[i];
    }

    for (int i = 0; assigned < nphoton; i = (i + 1) % ndev) {
        if (full <= 0.0 || (workload[i] > 0.f)) {
            share[i]++;
            assigned++;
        }
    }
}

extern "C" int mcxb_run_simulation_multi(const mcxb_config* cfg, const int* devices, int ndev, const float* workload, mcxb_output* out,
        mcxb_multi_info* info) {
    if (!cfg || !devices || !out || ndev < 1 || ndev > MCXB_MAX_DEVICES) {
        return fail(MCXB_ERR_ARG, "mcxb_run_simulation_multi: bad arguments (1..%d devices)", MCXB_MAX_DEVICES);
    }

    if (ndev == 1) {
        const int rc1 = mcxb_run_simulation(cfg, devices[0], out);

        if (info && rc1 == MCXB_OK) {
            memset(info, 0, sizeof(*info));
            info->ndev = 1;
            info->share[0] = cfg->nphoton;
            info->detected[0] = out->detected;
            info->kernel_ms[0] = out->runtime_ms;
            info->nthread[0] = out->nthread;
        }

        return rc1;
    }

    if (cfg->replay_seed) {
        return fail(MCXB_ERR_ARG, "replay should only work with a single device");       /* src/mcx_host.cpp:723 */
    }

    if (cfg->debuglevel & (MCXB_DEBUG_MOVE | MCXB_DEBUG_MOVE_ONLY | MCXB_DEBUG_RNG)) {
        return fail(MCXB_ERR_ARG, "trajectory capture and the RNG debug mode run on one device at a time");
    }

    for (int i = 0; i < ndev; i++) {
        for (int j = 0; j < i; j++) {
            if (devices[i] == devices[j]) {
                return fail(MCXB_ERR_ARG, "device %d is listed twice", devices[i]);
            }
        }
    }

    Nccl& N = nccl();
    const std::vector<int> devs(devices, devices + ndev);
    CommSet* cs = nullptr;
    int rc = MCXB_OK;
    static const bool timing = getenv("MCXB_TIMING") != nullptr;
    const double t_start = now_ms();
    double t_created = 0, t_kernels = 0, t_exchanged = 0;
    /* peer-memory exchange unless NCCL is asked for or a peer cannot be mapped (see the head of this file) */
    const char* how = getenv("MCXB_MULTI_EXCHANGE");
    {
        /* a first call creates one CUDA context per device (hundreds of ms each): all of them at once */
        std::vector<std::thread> th;

        for (int i = 0; i < ndev; i++) {
            th.emplace_back([&, i] {
                if (cudaSetDevice(devs[i]) == cudaSuccess) {
                    cudaFree(nullptr);
                }
            });
        }

        for (auto& t : th) {
            t.join();
        }
    }
    const double t_ctx = now_ms();
    const bool use_p2p = !(how && strcmp(how, "nccl") == 0) && enable_peers(devs);
    /* NCCL path: the communicators of a device list are built once per process (ncclCommInitAll: seconds for a first call);
     * a first call builds them on a helper thread while this one uploads the volumes and the photon kernels run */
    int comm_rc = MCXB_OK;
    std::string comm_err;
    std::thread comm_thread([&] {
        if (use_p2p) {
            return;
        }

        comm_rc = get_comms(devs, &cs);

        if (comm_rc != MCXB_OK) {
            comm_err = mcxb_last_error();
        }
    });

    std::vector<uint64_t> share(ndev);
    mcxb_split_photons(cfg->nphoton, workload, ndev, share.data());
    std::vector<mcxb_sim*> sims(ndev, (mcxb_sim*)nullptr);
    std::vector<int> rcs(ndev, MCXB_OK);
    std::vector<std::string> errs(ndev);
    std::vector<uint32_t> counts(ndev, 0), stored(ndev, 0), offs(ndev, 0);
    std::vector<float> ms(ndev, 0.f);
    uint32_t total = 0, reclen = 0, kept = 0;
    uint64_t fieldlen = 0, skip = 0;

    /* ---- one resident simulation per device, built concurrently (volume repack + H2D per device) ---- */
    {
        std::vector<std::thread> th;

        for (int i = 0; i < ndev; i++) {
            th.emplace_back([&, i] {
                mcxb_config c = *cfg;
                c.nphoton = share[i];
                c.isnormalized = (i == 0) ? cfg->isnormalized : 0;
                rcs[i] = mcxb_sim_create(&c, devs[i], &sims[i]);

                if (rcs[i] == MCXB_OK) {
                    rcs[i] = mcxb_sim_reset(sims[i], nullptr);
                }

                if (rcs[i] != MCXB_OK) {
                    errs[i] = mcxb_last_error();
                }
            });
        }

        for (auto& t : th) {
            t.join();
        }
    }

    for (int i = 0; i < ndev; i++) {
        if (rcs[i] != MCXB_OK) {
            rc = fail(rcs[i], "device %d: %s", devs[i], errs[i].c_str());
            goto done;
        }
    }

    t_created = now_ms();

    /* ---- seed slices: device i continues the ONE rand() stream where device i-1 stopped (src/mcx_host.cpp:759-768) ---- */
    for (int i = 0; i < ndev; i++) {
        if (i > 0) {
            rc = mcxb_sim_reseed(sims[i], cfg->seed, skip);

            if (rc != MCXB_OK) {
                goto done;
            }
        }

        skip += mcxb_sim_nthread(sims[i]);
    }

    /* ---- launch everywhere, then finalize (accumulators -> float32 volume) on the same streams ---- */
    if (cfg->respin > 1) {
        /* `-r R`: every device runs its share in R batches; batch k of device i takes slice k * (threads of all devices) +
         * (threads of devices 0..i-1) of the seed stream.  One host thread per device: a batch waits for the previous one */
        std::vector<std::thread> th;
        std::vector<float> bms(ndev, 0.f);
        uint64_t pre = 0;

        for (int i = 0; i < ndev; i++) {
            const uint64_t myskip = pre;
            th.emplace_back([&, i, myskip] {
                rcs[i] = mcxb_sim_run_batches(sims[i], share[i], (uint32_t)cfg->respin, cfg->seed, myskip, skip, &bms[i]);

                if (rcs[i] != MCXB_OK) {
                    errs[i] = mcxb_last_error();
                }
            });
            pre += mcxb_sim_nthread(sims[i]);
        }

        for (auto& t : th) {
            t.join();
        }

        for (int i = 0; i < ndev; i++) {
            if (rcs[i] != MCXB_OK) {
                rc = fail(rcs[i], "device %d: %s", devs[i], errs[i].c_str());
                goto done;
            }

            ms[i] = bms[i];
        }
    } else {
        for (int i = 0; i < ndev && rc == MCXB_OK; i++) {
            rc = mcxb_sim_launch(sims[i], nullptr);
        }
    }

    for (int i = 0; i < ndev && rc == MCXB_OK; i++) {
        rc = mcxb_sim_finalize(sims[i], nullptr);
    }

    if (rc != MCXB_OK) {
        goto done;
    }

    for (int i = 0; i < ndev; i++) {
        if (cfg->respin <= 1) {
            ms[i] = mcxb_sim_last_kernel_ms(sims[i]);   /* waits for the photon kernel of device i */
        }

        CUDA_TRY(cudaSetDevice(devs[i]));
        CUDA_TRY(cudaDeviceSynchronize());               /* finalize done: the exchange streams below are non-blocking ones */
    }

    t_kernels = now_ms();

    /* ---- the exchange step, over peer memory ---- */
    if (use_p2p) {
        comm_thread.join();
        fieldlen = mcxb_sim_fieldlen(sims[0]);
        reclen = mcxb_sim_reclen(sims[0]);
        CUDA_TRY(cudaSetDevice(devs[0]));

        if (cfg->issave2pt) {
            PeerList peers;
            peers.n = ndev - 1;

            for (int i = 1; i < ndev; i++) {
                peers.src[i - 1] = static_cast<const float*>(mcxb_sim_field_devptr(sims[i]));
            }

            const int grid = (int)std::min<uint64_t>((fieldlen / 4 + 255) / 256 + 1, 148 * 8);
            peer_sum_kernel <<< grid, 256>>>(static_cast<float*>(mcxb_sim_field_devptr(sims[0])), peers, (size_t)fieldlen);
            CUDA_TRY(cudaGetLastError());
        }

        {
            /* energy pairs and detected counts: a few bytes per device through the host */
            double esum[2] = { 0.0, 0.0 };

            for (int i = 0; i < ndev; i++) {
                double e[2];
                CUDA_TRY(cudaMemcpy(e, mcxb_sim_energy_devptr(sims[i]), sizeof(e), cudaMemcpyDeviceToHost));
                esum[0] += e[0];
                esum[1] += e[1];
                CUDA_TRY(cudaMemcpy(&counts[i], mcxb_sim_detcount_devptr(sims[i]), sizeof(uint32_t), cudaMemcpyDeviceToHost));
            }

            CUDA_TRY(cudaMemcpy(mcxb_sim_energy_devptr(sims[0]), esum, sizeof(esum), cudaMemcpyHostToDevice));
        }

        if (mcxb_sim_detphoton_devptr(sims[0])) {
            for (int i = 0; i < ndev; i++) {
                const uint32_t have = std::min(counts[i], cfg->maxdetphoton);
                offs[i] = kept;
                stored[i] = std::min(have, cfg->maxdetphoton - kept);
                kept += stored[i];
                total += counts[i];
            }

            for (int i = 1; i < ndev; i++) {
                if (!stored[i]) {
                    continue;
                }

                if (reclen) {
                    CUDA_TRY(cudaMemcpyPeerAsync((float*)mcxb_sim_detphoton_devptr(sims[0]) + (size_t)offs[i] * reclen, devs[0],
                                                 mcxb_sim_detphoton_devptr(sims[i]), devs[i], sizeof(float) * (size_t)stored[i] * reclen, 0));
                }

                if (mcxb_sim_seeddata_devptr(sims[0])) {
                    CUDA_TRY(cudaMemcpyPeerAsync((uint64_t*)mcxb_sim_seeddata_devptr(sims[0]) + (size_t)offs[i] * 2, devs[0],
                                                 mcxb_sim_seeddata_devptr(sims[i]), devs[i], 16 * (size_t)stored[i], 0));
                }
            }

            CUDA_TRY(cudaMemcpy(mcxb_sim_detcount_devptr(sims[0]), &total, sizeof(uint32_t), cudaMemcpyHostToDevice));
        }

        CUDA_TRY(cudaDeviceSynchronize());
        goto exchanged;
    }

    /* ---- the exchange step, over NCCL ---- */
    comm_thread.join();

    if (comm_rc != MCXB_OK) {
        rc = fail(comm_rc, "%s", comm_err.c_str());
        goto done;
    }

    fieldlen = mcxb_sim_fieldlen(sims[0]);
    reclen = mcxb_sim_reclen(sims[0]);
    NCCL_TRY(N.GroupStart());

    for (int i = 0; i < ndev; i++) {
        if (cfg->issave2pt) {
            NCCL_TRY(N.Reduce(mcxb_sim_field_devptr(sims[i]), mcxb_sim_field_devptr(sims[i]), fieldlen, ncclFloat, ncclSum, 0, cs->comm[i], cs->stream[i]));
        }

        NCCL_TRY(N.Reduce(mcxb_sim_energy_devptr(sims[i]), mcxb_sim_energy_devptr(sims[i]), 2, ncclDouble, ncclSum, 0, cs->comm[i], cs->stream[i]));
        NCCL_TRY(N.AllGather(mcxb_sim_detcount_devptr(sims[i]), cs->counts[i], 1, ncclUint32, cs->comm[i], cs->stream[i]));
    }

    NCCL_TRY(N.GroupEnd());
    CUDA_TRY(cudaSetDevice(devs[0]));
    CUDA_TRY(cudaMemcpyAsync(counts.data(), cs->counts[0], sizeof(uint32_t) * ndev, cudaMemcpyDeviceToHost, cs->stream[0]));
    CUDA_TRY(cudaStreamSynchronize(cs->stream[0]));

    if (mcxb_sim_detphoton_devptr(sims[0])) {
        /* exclusive scan, clipped at the buffer size the way the reference clips at maxdetphoton (:1207-1216) */
        for (int i = 0; i < ndev; i++) {
            const uint32_t have = std::min(counts[i], cfg->maxdetphoton);
            offs[i] = kept;
            stored[i] = std::min(have, cfg->maxdetphoton - kept);
            kept += stored[i];
            total += counts[i];
        }

        NCCL_TRY(N.GroupStart());

        for (int i = 1; i < ndev; i++) {
            if (!stored[i]) {
                continue;
            }

            if (reclen) {
                NCCL_TRY(N.Send(mcxb_sim_detphoton_devptr(sims[i]), (size_t)stored[i] * reclen, ncclFloat, 0, cs->comm[i], cs->stream[i]));
                NCCL_TRY(N.Recv((float*)mcxb_sim_detphoton_devptr(sims[0]) + (size_t)offs[i] * reclen, (size_t)stored[i] * reclen, ncclFloat, i, cs->comm[0], cs->stream[0]));
            }

            if (mcxb_sim_seeddata_devptr(sims[0])) {
                NCCL_TRY(N.Send(mcxb_sim_seeddata_devptr(sims[i]), (size_t)stored[i] * 2, ncclUint64, 0, cs->comm[i], cs->stream[i]));
                NCCL_TRY(N.Recv((uint64_t*)mcxb_sim_seeddata_devptr(sims[0]) + (size_t)offs[i] * 2, (size_t)stored[i] * 2, ncclUint64, i, cs->comm[0], cs->stream[0]));
            }
        }

        NCCL_TRY(N.GroupEnd());
        /* device 0 now holds every record: its counter becomes the whole job's, which is what fetch reports and copies */
        CUDA_TRY(cudaMemcpyAsync(mcxb_sim_detcount_devptr(sims[0]), &total, sizeof(uint32_t), cudaMemcpyHostToDevice, cs->stream[0]));
    }

    for (int i = 0; i < ndev; i++) {
        CUDA_TRY(cudaSetDevice(devs[i]));
        CUDA_TRY(cudaStreamSynchronize(cs->stream[i]));
    }

exchanged:
    t_exchanged = now_ms();
    /* ---- one read-back, one normalisation with the global launched energy (src/mcx_host.cpp:1382-1465) ---- */
    rc = mcxb_sim_fetch(sims[0], nullptr, out);

    if (timing) {
        fprintf(stderr, "mcxb multi (%s, %d devices): contexts %.1f ms, create %.1f ms, kernels %.1f ms, exchange %.1f ms, fetch %.1f ms\n",
                use_p2p ? "peer memory" : "nccl", ndev, t_ctx - t_start, t_created - t_ctx, t_kernels - t_created, t_exchanged - t_kernels, now_ms() - t_exchanged);
    }

    if (rc == MCXB_OK) {
        out->runtime_ms = *std::max_element(ms.begin(), ms.end());
        out->kernel_launches = 2 * (uint64_t)ndev;
        out->saved = std::min(out->saved, kept);

        if (info) {
            memset(info, 0, sizeof(*info));
            info->ndev = ndev;
            info->nccl_version = use_p2p ? 0 : mcxb_nccl_version();      /* 0 = the exchange went over peer memory */

            for (int i = 0; i < ndev; i++) {
                info->share[i] = share[i];
                info->detected[i] = counts[i];
                info->kernel_ms[i] = ms[i];
                info->nthread[i] = mcxb_sim_nthread(sims[i]);
            }
        }
    }

done:

    if (comm_thread.joinable()) {
        comm_thread.join();
    }

    {
        const double t0 = now_ms();

        for (int i = 0; i < ndev; i++) {
            mcxb_sim_destroy(sims[i]);
        }

        if (timing) {
            fprintf(stderr, "mcxb multi: destroy %.1f ms, whole call %.1f ms\n", now_ms() - t0, now_ms() - t_start);
        }
    }

    return rc;
}
