/*
 * testhooks.cu -- unit-level entry points of the C ABI (include/mcxb200.h, "unit-level hooks").
 * Each one launches a tiny kernel built from the SAME device functions as photon_kernel
 * (photon_device.cuh), so the bit-exact tier is tested on the code that ships.
 */
#include "../../include/mcxb200.h"
#include "photon_device.cuh"
#include <vector>
#include <string>

using namespace mcxb;

#define HOOK_TRY(call)                         \
    do {                                       \
        cudaError_t e__ = (call);              \
        if (e__ != cudaSuccess) {              \
            rc = MCXB_ERR_CUDA_BASE - (int)e__; \
            goto done;                         \
        }                                      \
    } while (0)

__global__ void hook_rng(const uint32_t* __restrict__ seeds, uint32_t n, uint32_t ndraw, float* __restrict__ out, unsigned long long* __restrict__ state) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

    if (i >= n) {
        return;
    }

    Rng r;
    rng_seed(r, seeds + 4 * (size_t)i);

    for (uint32_t k = 0; k < ndraw; k++) {
        out[(size_t)i * ndraw + k] = rng_uniform(r);
    }

    state[2 * (size_t)i] = r.a;
    state[2 * (size_t)i + 1] = r.b;
}

__global__ void hook_trace(const float4* __restrict__ p0, const float4* __restrict__ v0, uint32_t n, uint32_t nstep,
                           uint32_t nx, uint32_t ny, uint32_t nz, float musp, mcxb_trace_step* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

    if (i >= n) {
        return;
    }

    float px = p0[i].x, py = p0[i].y, pz = p0[i].z;
    const float vx = v0[i].x, vy = v0[i].y, vz = v0[i].z;
    int ix = (int)(short)floorf(px), iy = (int)(short)floorf(py), iz = (int)(short)floorf(pz), face = -1;
    mcxb_trace_step last;
    bool done = false;

    for (uint32_t k = 0; k < nstep; k++) {
        if (!done) {
            /* the stepping statements of photon_kernel with an infinite scattering length */
            const float dist = face_distance(px, py, pz, vx, vy, vz, ix, iy, iz, face);
            float slen;
            const float len = step_length(dist, musp, __int_as_float(0x7F800000), slen);
            px = advance(px, len, vx);
            py = advance(py, len, vy);
            pz = advance(pz, len, vz);

            if (face == 0) {
                ix += (vx > 0.f ? 1 : -1);
            } else if (face == 1) {
                iy += (vy > 0.f ? 1 : -1);
            } else {
                iz += (vz > 0.f ? 1 : -1);
            }

            last.dist = len;
            last.px = px;
            last.py = py;
            last.pz = pz;
            last.ix = (int16_t)ix;
            last.iy = (int16_t)iy;
            last.iz = (int16_t)iz;
            last.face = (int16_t)face;

            if ((uint32_t)(ix & 0xFFFF) >= nx || (uint32_t)(iy & 0xFFFF) >= ny || (uint32_t)(iz & 0xFFFF) >= nz) {
                last.idx1d = (ix < 0 || iy < 0 || iz < 0) ? kOutsideMin : kOutsideMax;
                done = true;
            } else {
                last.idx1d = (uint32_t)(iz * (int)(nx * ny) + iy * (int)nx + ix);
            }
        }

        out[(size_t)i * nstep + k] = last;
    }
}

__global__ void hook_scalar(const float* __restrict__ a, const int* __restrict__ dir, uint32_t n, float* __restrict__ na,
                            const float4* __restrict__ v, const float* __restrict__ n1, const float* __restrict__ n2,
                            const int* __restrict__ face, uint32_t m, float* __restrict__ rc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

    if (i < n) {
        na[i] = nudge(a[i], dir[i]);
    }

    if (i < m) {
        rc[i] = fresnel(v[i].x, v[i].y, v[i].z, n1[i], n2[i], face[i]);
    }
}

__global__ void hook_rotate(float4* __restrict__ v, const float* st, const float* ct, const float* sp, const float* cp, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

    if (i < n) {
        float4 t = v[i];
        rotate_direction(t.x, t.y, t.z, st[i], ct[i], sp[i], cp[i]);
        v[i] = t;
    }
}

__global__ void hook_refract(float4* __restrict__ v, const float* n1, const float* n2, const int* face, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;

    if (i < n) {
        float4 t = v[i];
        refract(t.x, t.y, t.z, n1[i], n2[i], face[i]);
        v[i] = t;
    }
}

/* microbenchmark used by bench.py / profiles to measure the L2 reduction ceiling this GPU offers:
 * every thread issues `iters` fire-and-forget reductions to pseudo-random addresses of a `span`-element
 * buffer (element = 4 or 8 bytes). */
template <typename T>
__global__ void hook_redbench(T* __restrict__ buf, uint32_t span, uint32_t iters, uint32_t hot_permille) {
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;

    for (uint32_t k = 0; k < iters; k++) {
        x ^= x << 13;
        x ^= x >> 17;
        x ^= x << 5;
        uint32_t idx = x % span;

        if ((x >> 22) % 1000u < hot_permille) {
            idx = idx & 63u;     /* a clustered "near the source" address */
        }

        red_add(buf + idx, 1.0f);
    }
}

namespace {
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() {
        cudaFree(p);
    }
    cudaError_t alloc(size_t n) {
        return cudaMalloc(&p, n ? n : 1);
    }
};
}

extern "C" int mcxb_test_rng(int device, const uint32_t* seeds, uint32_t n, uint32_t ndraw, float* out, uint64_t* state_out) {
    int rc = MCXB_OK;
    DevBuf ds, dout, dst;
    HOOK_TRY(cudaSetDevice(device));
    HOOK_TRY(ds.alloc(16 * (size_t)n));
    HOOK_TRY(dout.alloc(4 * (size_t)n * ndraw));
    HOOK_TRY(dst.alloc(16 * (size_t)n));
    HOOK_TRY(cudaMemcpy(ds.p, seeds, 16 * (size_t)n, cudaMemcpyHostToDevice));
    hook_rng <<< (n + 127) / 128, 128>>>((const uint32_t*)ds.p, n, ndraw, (float*)dout.p, (unsigned long long*)dst.p);
    HOOK_TRY(cudaGetLastError());
    HOOK_TRY(cudaMemcpy(out, dout.p, 4 * (size_t)n * ndraw, cudaMemcpyDeviceToHost));

    if (state_out) {
        HOOK_TRY(cudaMemcpy(state_out, dst.p, 16 * (size_t)n, cudaMemcpyDeviceToHost));
    }

done:
    return rc;
}

extern "C" int mcxb_test_trace(int device, const mcxb_f4* p0, const mcxb_f4* v0, uint32_t n, uint32_t nstep,
                               uint32_t dimx, uint32_t dimy, uint32_t dimz, float musp, mcxb_trace_step* out) {
    int rc = MCXB_OK;
    DevBuf dp, dv, dout;
    HOOK_TRY(cudaSetDevice(device));
    HOOK_TRY(dp.alloc(16 * (size_t)n));
    HOOK_TRY(dv.alloc(16 * (size_t)n));
    HOOK_TRY(dout.alloc(sizeof(mcxb_trace_step) * (size_t)n * nstep));
    HOOK_TRY(cudaMemcpy(dp.p, p0, 16 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(dv.p, v0, 16 * (size_t)n, cudaMemcpyHostToDevice));
    hook_trace <<< (n + 127) / 128, 128>>>((const float4*)dp.p, (const float4*)dv.p, n, nstep, dimx, dimy, dimz, musp, (mcxb_trace_step*)dout.p);
    HOOK_TRY(cudaGetLastError());
    HOOK_TRY(cudaMemcpy(out, dout.p, sizeof(mcxb_trace_step) * (size_t)n * nstep, cudaMemcpyDeviceToHost));
done:
    return rc;
}

extern "C" int mcxb_test_scalar(int device, const float* a, const int32_t* dir, uint32_t n, float* nextafter_out,
                                const mcxb_f4* v, const float* n1, const float* n2, const int32_t* face, uint32_t m, float* rcoef_out) {
    int rc = MCXB_OK;
    DevBuf da, dd, dna, dv, d1, d2, df, drc;
    const uint32_t nn = n > m ? n : m;
    HOOK_TRY(cudaSetDevice(device));
    HOOK_TRY(da.alloc(4 * (size_t)n));
    HOOK_TRY(dd.alloc(4 * (size_t)n));
    HOOK_TRY(dna.alloc(4 * (size_t)n));
    HOOK_TRY(dv.alloc(16 * (size_t)m));
    HOOK_TRY(d1.alloc(4 * (size_t)m));
    HOOK_TRY(d2.alloc(4 * (size_t)m));
    HOOK_TRY(df.alloc(4 * (size_t)m));
    HOOK_TRY(drc.alloc(4 * (size_t)m));
    HOOK_TRY(cudaMemcpy(da.p, a, 4 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(dd.p, dir, 4 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(dv.p, v, 16 * (size_t)m, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(d1.p, n1, 4 * (size_t)m, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(d2.p, n2, 4 * (size_t)m, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(df.p, face, 4 * (size_t)m, cudaMemcpyHostToDevice));

    if (nn) {
        hook_scalar <<< (nn + 127) / 128, 128>>>((const float*)da.p, (const int*)dd.p, n, (float*)dna.p, (const float4*)dv.p,
                (const float*)d1.p, (const float*)d2.p, (const int*)df.p, m, (float*)drc.p);
        HOOK_TRY(cudaGetLastError());
    }

    HOOK_TRY(cudaMemcpy(nextafter_out, dna.p, 4 * (size_t)n, cudaMemcpyDeviceToHost));
    HOOK_TRY(cudaMemcpy(rcoef_out, drc.p, 4 * (size_t)m, cudaMemcpyDeviceToHost));
done:
    return rc;
}

extern "C" int mcxb_test_rotate(int device, mcxb_f4* v, const float* st, const float* ct, const float* sp, const float* cp, uint32_t n) {
    int rc = MCXB_OK;
    DevBuf dv, a, b, c, d;
    HOOK_TRY(cudaSetDevice(device));
    HOOK_TRY(dv.alloc(16 * (size_t)n));
    HOOK_TRY(a.alloc(4 * (size_t)n));
    HOOK_TRY(b.alloc(4 * (size_t)n));
    HOOK_TRY(c.alloc(4 * (size_t)n));
    HOOK_TRY(d.alloc(4 * (size_t)n));
    HOOK_TRY(cudaMemcpy(dv.p, v, 16 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(a.p, st, 4 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(b.p, ct, 4 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(c.p, sp, 4 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(d.p, cp, 4 * (size_t)n, cudaMemcpyHostToDevice));
    hook_rotate <<< (n + 127) / 128, 128>>>((float4*)dv.p, (const float*)a.p, (const float*)b.p, (const float*)c.p, (const float*)d.p, n);
    HOOK_TRY(cudaGetLastError());
    HOOK_TRY(cudaMemcpy(v, dv.p, 16 * (size_t)n, cudaMemcpyDeviceToHost));
done:
    return rc;
}

extern "C" int mcxb_test_refract(int device, mcxb_f4* v, const float* n1, const float* n2, const int32_t* face, uint32_t n) {
    int rc = MCXB_OK;
    DevBuf dv, a, b, c;
    HOOK_TRY(cudaSetDevice(device));
    HOOK_TRY(dv.alloc(16 * (size_t)n));
    HOOK_TRY(a.alloc(4 * (size_t)n));
    HOOK_TRY(b.alloc(4 * (size_t)n));
    HOOK_TRY(c.alloc(4 * (size_t)n));
    HOOK_TRY(cudaMemcpy(dv.p, v, 16 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(a.p, n1, 4 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(b.p, n2, 4 * (size_t)n, cudaMemcpyHostToDevice));
    HOOK_TRY(cudaMemcpy(c.p, face, 4 * (size_t)n, cudaMemcpyHostToDevice));
    hook_refract <<< (n + 127) / 128, 128>>>((float4*)dv.p, (const float*)a.p, (const float*)b.p, (const int*)c.p, n);
    HOOK_TRY(cudaGetLastError());
    HOOK_TRY(cudaMemcpy(v, dv.p, 16 * (size_t)n, cudaMemcpyDeviceToHost));
done:
    return rc;
}

extern "C" int mcxb_bench_red(int device, int elem_bytes, uint64_t span_elems, uint32_t nblock, uint32_t iters,
                              uint32_t hot_permille, uint32_t repeats, float* ms_out, uint64_t* ops_out) {
    int rc = MCXB_OK;
    DevBuf buf;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    HOOK_TRY(cudaSetDevice(device));

    if ((elem_bytes != 4 && elem_bytes != 8) || span_elems == 0 || span_elems > 0xFFFFFFFFull) {
        return MCXB_ERR_ARG;
    }

    HOOK_TRY(buf.alloc((size_t)elem_bytes * span_elems));
    HOOK_TRY(cudaMemset(buf.p, 0, (size_t)elem_bytes * span_elems));
    HOOK_TRY(cudaEventCreate(&e0));
    HOOK_TRY(cudaEventCreate(&e1));

    for (uint32_t r = 0; r < repeats + 1; r++) {
        if (r == 1) {
            HOOK_TRY(cudaEventRecord(e0));
        }

        if (elem_bytes == 4) {
            hook_redbench<float> <<< nblock, 256>>>((float*)buf.p, (uint32_t)span_elems, iters, hot_permille);
        } else {
            hook_redbench<double> <<< nblock, 256>>>((double*)buf.p, (uint32_t)span_elems, iters, hot_permille);
        }
    }

    HOOK_TRY(cudaGetLastError());
    HOOK_TRY(cudaEventRecord(e1));
    HOOK_TRY(cudaEventSynchronize(e1));
    HOOK_TRY(cudaEventElapsedTime(ms_out, e0, e1));
    *ms_out /= (float)repeats;
    *ops_out = (uint64_t)nblock * 256ull * iters;
done:

    if (e0) {
        cudaEventDestroy(e0);
    }

    if (e1) {
        cudaEventDestroy(e1);
    }

    return rc;
}


/* ---------------------------------------------------------------------------------------------------
 * Where does a reduction execute?  (profiles/r2_deposit_variants.md: "50 % of the RED sectors miss in the L2 lookup with
 * no DRAM traffic".)  Only the SMs selected by `mode` issue `iters` fp64 reductions per thread, all of them into ONE
 * 32-byte sector: mode 0 = every SM, 1 / 2 = lower / upper half of the SM ids, 3 / 4 = even / odd SM ids.  Run under
 * `ncu --metrics lts__t_sectors_srcunit_tex_op_red_lookup_{hit,miss}.sum`: if the lookup-miss count follows WHICH SMs
 * issue the reductions, the "miss" is the hop from the issuing die's L2 to the L2 partition that owns the line, not a
 * cache miss.
 * ------------------------------------------------------------------------------------------------- */
__global__ void hook_red_die(double* buf, uint32_t iters, int mode, uint32_t nsm, unsigned long long* issued) {
    uint32_t smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    const bool on = mode == 0 || (mode == 1 && smid < nsm / 2) || (mode == 2 && smid >= nsm / 2) || (mode == 3 && (smid & 1u) == 0) || (mode == 4 && (smid & 1u) == 1);

    if (!on) {
        return;
    }

    for (uint32_t i = 0; i < iters; i++) {
        red_add(buf + (threadIdx.x & 3u), 1.f);
    }

    if (threadIdx.x == 0) {
        atomicAdd(issued, (unsigned long long)blockDim.x * iters);
    }
}

extern "C" int mcxb_bench_red_die(int device, int mode, uint32_t nblock, uint32_t iters, uint64_t* issued_out, double* sum_out) {
    int rc = MCXB_OK;
    DevBuf buf, cnt;
    int nsm = 0;
    double h[4] = {0, 0, 0, 0};
    HOOK_TRY(cudaSetDevice(device));
    HOOK_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device));
    HOOK_TRY(buf.alloc(256));
    HOOK_TRY(cnt.alloc(8));
    HOOK_TRY(cudaMemset(buf.p, 0, 256));
    HOOK_TRY(cudaMemset(cnt.p, 0, 8));
    hook_red_die <<< nblock, 256>>>((double*)buf.p, iters, mode, (uint32_t)nsm, (unsigned long long*)cnt.p);
    HOOK_TRY(cudaGetLastError());
    HOOK_TRY(cudaDeviceSynchronize());
    HOOK_TRY(cudaMemcpy(issued_out, cnt.p, 8, cudaMemcpyDeviceToHost));
    HOOK_TRY(cudaMemcpy(h, buf.p, 32, cudaMemcpyDeviceToHost));
    *sum_out = h[0] + h[1] + h[2] + h[3];
done:
    return rc;
}
