/*
 * CL/cl.h stand-in for building the reference's front-ends against the CUDA boundary.
 *
 * The reference's boundary header (src/mcx_host.h:40-44) includes <CL/cl.h> and spells its scalar and
 * vector members with OpenCL host typedefs (cl_uint, cl_float4, ...) and its device handles as
 * cl_platform_id / cl_device_id (src/mcx_host.h:168-169; src/pmcxcl.cpp:1132; src/mcxlabcl.cpp:106-107).
 * Nothing above the boundary calls an OpenCL function, so plain typedefs are all that is needed to
 * compile mcxcl.c / pmcxcl.cpp / mcxlabcl.cpp unchanged without an OpenCL SDK.  Handles are opaque:
 * integration/mcx_cuda_host.cpp stores (CUDA device ordinal + 1) in a cl_device_id.
 */
#ifndef MCXB200_CL_STUB_H
#define MCXB200_CL_STUB_H

#include <stdint.h>

typedef int8_t   cl_char;
typedef uint8_t  cl_uchar;
typedef int16_t  cl_short;
typedef uint16_t cl_ushort;
typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef int64_t  cl_long;
typedef uint64_t cl_ulong;
typedef float    cl_float;
typedef double   cl_double;
typedef cl_ulong cl_bitfield;

typedef union { cl_float s[4]; struct { cl_float x, y, z, w; }; } __attribute__((aligned(16))) cl_float4;
typedef union { cl_uint  s[4]; struct { cl_uint  x, y, z, w; }; } __attribute__((aligned(16))) cl_uint4;
typedef union { cl_uint  s[2]; struct { cl_uint  x, y; }; }       __attribute__((aligned(8)))  cl_uint2;

typedef struct mcxb_cl_platform* cl_platform_id;
typedef struct mcxb_cl_device*   cl_device_id;
typedef struct mcxb_cl_event*    cl_event;

#define CL_SUCCESS 0
#define CL_MEM_READ_ONLY      (1 << 2)
#define CL_MEM_WRITE_ONLY     (1 << 1)
#define CL_MEM_READ_WRITE     (1 << 0)
#define CL_MEM_COPY_HOST_PTR  (1 << 5)
#define CL_MEM_ALLOC_HOST_PTR (1 << 4)

#endif
