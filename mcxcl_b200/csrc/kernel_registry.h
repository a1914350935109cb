/*
 * kernel_registry.h -- table of the ahead-of-time specialisations of photon_kernel.
 *
 * The reference JIT-compiles one OpenCL program per run with -D flags for source type, reflection and
 * detector capture (src/mcx_host.cpp:857-971).  Here the same axes are C++ template parameters and
 * every combination that is shipped is compiled for sm_100a at build time; kernels_inst.cu is compiled
 * once per group (-DMCXB_INST_GROUP=k) so the groups build in parallel.
 */
#pragma once
#include "photon_kernel.cuh"

namespace mcxb {

typedef void (*PhotonKernelFn)(const SimParam);

struct KernelEntry {
    int  src;          /* SrcType or srcAny */
    bool reflect;
    int  savedet;      /* 0 none, 1 default record folded at compile time, 2 record flags at run time */
    bool media16, media32, acc64, stats, generic;     /* media word: 8 bits, 16 bits, or 32 bits (continuous media) */
    int  queue;        /* depth of the scattering queue (shared memory, 16 bytes x depth per thread), 0 = none */
    bool bcodes;       /* common kernel that honours per-face boundary codes / detect-on-face flags (`-B`, `--bc`) */
    bool ext;          /* extended physics compiled in (polarised light, RF, split voxels, adjoint detector sources): generic kernels with the extra per-packet state */
    PhotonKernelFn fn;
    const char* name;
};

constexpr int kNumGroups = 12;

} // namespace mcxb

extern "C" const mcxb::KernelEntry* mcxb_kernel_group_0(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_1(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_2(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_3(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_4(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_5(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_6(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_7(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_8(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_9(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_10(int* n);
extern "C" const mcxb::KernelEntry* mcxb_kernel_group_11(int* n);
