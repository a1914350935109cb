/*
 * clshim.h -- a minimal OpenCL-C 1.2 emulation layer for g++ (TEST INFRASTRUCTURE ONLY).
 *
 * Purpose: lets the reference's own kernel source (/root/reference/src/mcx_core.cl, OpenCL
 * branch) be compiled as ordinary host C++ so that it can serve as the strongest available
 * oracle for the B200 kernel.  Nothing here is part of the product path;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load the library built from it (oracle/_ref/libmcxref.so).
 *
 * Numeric contract implemented by this shim (the "bit-exact tier" of DESIGN.md):
 *   - IEEE-754 binary32 everywhere, no FMA contraction (build with -ffp-contract=off),
 *   - native_divide(a,b) == a/b (round-to-nearest), native_{sin,cos,log,exp,sqrt} == libm float,
 *   - rsqrt(x) == 1.f/sqrtf(x).
 */
#ifndef MCXB200_ORACLE_CLSHIM_H
#define MCXB200_ORACLE_CLSHIM_H

#include <sys/types.h>   /* uint, ushort, ulong (64-bit on LP64) */
#include <cmath>
#include <cfloat>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

using std::min;
using std::max;
using std::isnan;
using std::isinf;
/* pick the float overloads, as OpenCL C does for float arguments */
using std::fabs;
using std::floor;
using std::rint;
using std::acos;
using std::fmin;
using std::fmax;
using std::sqrt;

/* ---- address-space / kernel qualifiers vanish on the host ---- */
#define __global
#define __local
#define __private
#define __kernel
#define __constant const

/* ---- cl_khr_fp16 as far as updateproperty needs it (MED_TYPE 99 / 100 / 102, src/mcx_core.cl:1103-1152): a `half`
 *      is its 16 storage bits, vload_half widens them exactly to binary32, convert_float is the identity ---- */
typedef unsigned short half;
static inline float vload_half(size_t offset, const half* p) {
    const unsigned int h = p[offset];
    const unsigned int sign = (h & 0x8000u) << 16;
    unsigned int e = (h >> 10) & 0x1Fu, m = h & 0x3FFu, bits;

    if (e == 0) {
        if (m == 0) {
            bits = sign;
        } else {                /* subnormal half: normalise */
            e = 113;

            while (!(m & 0x400u)) {
                m <<= 1;
                e--;
            }

            bits = sign | (e << 23) | ((m & 0x3FFu) << 13);
        }
    } else if (e == 31) {
        bits = sign | 0x7F800000u | (m << 13);
    } else {
        bits = sign | ((e + 112) << 23) | (m << 13);
    }

    float f;
    memcpy(&f, &bits, 4);
    return f;
}
static inline float convert_float(float v) {
    return v;
}

/* ---- vector types (only the members/operators mcx_core.cl actually uses) ---- */
struct alignas(16) float4 {
    float x, y, z, w;
    float4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    explicit float4(float s) : x(s), y(s), z(s), w(s) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
};
static inline float4 operator+(float4 a, float4 b) {
    return float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
static inline float4 operator-(float4 a, float4 b) {
    return float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
static inline float4 operator*(float4 a, float4 b) {
    return float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w);
}
static inline float4 operator*(float4 a, float s) {
    return float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
static inline float4 operator*(float s, float4 a) {
    return a * s;
}
static inline float4 operator/(float4 a, float4 b) {
    return float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w);
}

/* OpenCL float3 occupies 16 bytes */
struct alignas(16) float3 {
    float x, y, z;
    float3() : x(0.f), y(0.f), z(0.f) {}
    explicit float3(float s) : x(s), y(s), z(s) {}
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
};
static inline float3 operator+(float3 a, float3 b) {
    return float3(a.x + b.x, a.y + b.y, a.z + b.z);
}
static inline float3 operator-(float3 a, float3 b) {
    return float3(a.x - b.x, a.y - b.y, a.z - b.z);
}
static inline float3 operator-(float3 a) {
    return float3(-a.x, -a.y, -a.z);
}
static inline float3 operator*(float3 a, float3 b) {
    return float3(a.x * b.x, a.y * b.y, a.z * b.z);
}
static inline float3 operator*(float3 a, float s) {
    return float3(a.x * s, a.y * s, a.z * s);
}
static inline float3 operator*(float s, float3 a) {
    return a * s;
}
static inline float3& operator+=(float3& a, float3 b) {
    a.x += b.x;
    a.y += b.y;
    a.z += b.z;
    return a;
}
static inline float3& operator*=(float3& a, float s) {
    a.x *= s;
    a.y *= s;
    a.z *= s;
    return a;
}
static inline float dot(float3 a, float3 b) {
    return a.x * b.x + a.y * b.y + a.z * b.z;
}

struct float2 {
    float x, y;
};
struct alignas(8) short4 {
    short x, y, z, w;
    short4() : x(0), y(0), z(0), w(0) {}
    short4(int a, int b, int c, int d) : x((short)a), y((short)b), z((short)c), w((short)d) {}
};
struct alignas(8) short3 {
    short x, y, z;
    short3() : x(0), y(0), z(0) {}
    short3(int a, int b, int c) : x((short)a), y((short)b), z((short)c) {}
};
struct alignas(16) int4 {
    int x, y, z, w;
    int4() : x(0), y(0), z(0), w(0) {}
    int4(int a, int b, int c, int d) : x(a), y(b), z(c), w(d) {}
};
struct alignas(16) uint4 {
    unsigned int x, y, z, w;
};
struct alignas(8) uint2 {
    unsigned int x, y;
};

/* per-host-thread work counters (segments / deposits / scatters per photon) */
extern thread_local unsigned long long clshim_cnt_isgreater, clshim_cnt_xchg, clshim_cnt_log;

/* ---- builtins ---- */
static inline float4 fabs(float4 a) {
    return float4(fabsf(a.x), fabsf(a.y), fabsf(a.z), fabsf(a.w));
}
/* OpenCL vector relational builtins return -1 (all bits set) for true */
static inline int4 isgreater(float4 a, float4 b) {
    clshim_cnt_isgreater++;   /* exactly one call per hitgrid() */
    return int4(-(a.x > b.x), -(a.y > b.y), -(a.z > b.z), -(a.w > b.w));
}
static inline float4 convert_float4_rtp(short4 a) {
    return float4((float)a.x, (float)a.y, (float)a.z, (float)a.w);
}
static inline float4 convert_float4_rtp(int4 a) {
    return float4((float)a.x, (float)a.y, (float)a.z, (float)a.w);
}
static inline short convert_short_rtn(float v) {
    return (short)floorf(v);
}
static inline short convert_short_rte(float v) {
    return (short)rintf(v);
}
static inline float convert_float_rte(float v) {
    return rintf(v);
}

static inline float  native_divide(float a, float b)   {
    return a / b;
}
static inline float4 native_divide(float4 a, float4 b) {
    return a / b;
}
static inline float native_sin(float x)  {
    return sinf(x);
}
static inline float native_cos(float x)  {
    return cosf(x);
}
static inline float native_log(float x)  {
    clshim_cnt_log++;
    return logf(x);
}
static inline float native_exp(float x)  {
    return expf(x);
}
static inline float native_sqrt(float x) {
    return sqrtf(x);
}
static inline float rsqrt(float x)       {
    return 1.f / sqrtf(x);
}
static inline float sincos(float x, float* c) {
    *c = cosf(x);
    return sinf(x);
}
static inline unsigned int as_uint(float f) {
    unsigned int u;
    memcpy(&u, &f, 4);
    return u;
}
static inline int as_int(float f) {
    int u;
    memcpy(&u, &f, 4);
    return u;
}

/* ---- atomics: one work-item runs at a time on a PRIVATE output buffer, so plain ops suffice ---- */
static inline unsigned int atomic_inc(volatile unsigned int* p) {
    return (*p)++;
}
static inline unsigned int atomic_dec(volatile unsigned int* p) {
    return (*p)--;
}
static inline int atomic_dec(volatile int* p) {
    return (*p)--;
}
static inline float atomic_xchg(volatile float* p, float v) {
    clshim_cnt_xchg++;        /* the CAS-fallback atomicadd issues two exchanges per add */
    float old = *p;
    *p = v;
    return old;
}

struct ClShimWorkItem {
    size_t global_id, local_id, local_size, num_groups, group_id;
};
extern thread_local ClShimWorkItem clshim_wi;

static inline size_t get_global_id(int)   {
    return clshim_wi.global_id;
}
static inline size_t get_local_id(int)    {
    return clshim_wi.local_id;
}
static inline size_t get_local_size(int)  {
    return clshim_wi.local_size;
}
static inline size_t get_num_groups(int)  {
    return clshim_wi.num_groups;
}
static inline size_t get_group_id(int)    {
    return clshim_wi.group_id;
}
static inline size_t get_global_size(int) {
    return clshim_wi.local_size * clshim_wi.num_groups;
}
#define CLK_LOCAL_MEM_FENCE 0
static inline void barrier(int) {}

#endif
