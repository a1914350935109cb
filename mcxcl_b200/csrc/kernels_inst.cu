/*
 * kernels_inst.cu -- explicit instantiations of photon_kernel, one group per compilation
 * (nvcc ... -DMCXB_INST_GROUP=k).  See kernel_registry.h.
 *
 *   "common" = the GEN=false form of the kernel (photon_kernel.cuh), "generic" = every option at run time; every
 *   common group also carries its fp64 kernels with the scattering queue (q8)
 *   group 0: pencil beam,           8-bit media, common
 *   group 1: disk source,           8-bit media, common
 *   group 2: planar + fourier,      8-bit media, common
 *   group 3: isotropic + cone,      8-bit media, common
 *   group 4: any source (run time), 8-bit media, common
 *   group 5: any source (run time), 8-bit media, generic  (+ the instrumented variant that counts segments/deposits/scatters)
 *   group 6: any source (run time), 16-bit media, generic (volumes with more than 127 labels)
 *   group 7: any source (run time), 32-bit media words, generic (continuous media formats 99-104: the word encodes the properties)
 *   group 8: any source (run time), 8-bit media, generic + extended physics (polarised light, RF forward / replay)
 *   group 9: any source (run time), 16- and 32-bit media, generic + extended physics
 *   group 10 / 11: pencil beam / any source, 8-bit media, common kernels that honour per-face boundary codes (`-B`, `--bc`)
 */
#include "kernel_registry.h"

#ifndef MCXB_INST_GROUP
    #error "compile with -DMCXB_INST_GROUP=<0..11>"
#endif

namespace mcxb {

#define MCXB_STR2(x) #x
#define MCXB_STR(x) MCXB_STR2(x)
#define MCXB_KQ(SRC, R, D, M, A, S, G, Q) { SRC, R, D, sizeof(M) == 2, sizeof(M) == 4, sizeof(A) == 8, S, G, Q, false, false, photon_kernel<SRC, R, D, M, A, S, G, Q>, #SRC "/" #R "/det" #D "/" #M "/" #A "/" #G "/q" MCXB_STR(Q) }
/* generic kernel with the extended physics, reflection x detector capture */
#define MCXB_KX(R, D, M) { srcAny, R, D, sizeof(M) == 2, sizeof(M) == 4, true, false, true, 0, false, true, photon_kernel<srcAny, R, D, M, double, false, true, 0, true>, "srcAny/" #R "/det" #D "/" #M "/double/true/ext" }
/* common kernels with per-face boundary codes (8-bit media, fp64), with and without the scattering queue */
#define MCXB_KB(SRC, R, D, Q) { SRC, R, D, false, false, true, false, false, Q, true, false, photon_kernel<SRC, R, D, uint8_t, double, false, false, Q, false, true>, #SRC "/" #R "/det" #D "/uint8_t/double/false/q" MCXB_STR(Q) "/bc" }
#define MCXB_RDB(SRC, Q) MCXB_KB(SRC, false, 0, Q), MCXB_KB(SRC, true, 0, Q), MCXB_KB(SRC, false, 1, Q), MCXB_KB(SRC, true, 1, Q)
#define MCXB_RDX(M) MCXB_KX(false, 0, M), MCXB_KX(true, 0, M), MCXB_KX(false, 1, M), MCXB_KX(true, 1, M)
#define MCXB_K(SRC, R, D, M, A, S, G) MCXB_KQ(SRC, R, D, M, A, S, G, 0)
/* reflection x detector capture (0 = none, 1 = default record) */
#define MCXB_RD(SRC, M, A, G) MCXB_K(SRC, false, 0, M, A, false, G), MCXB_K(SRC, true, 0, M, A, false, G), \
                              MCXB_K(SRC, false, 1, M, A, false, G), MCXB_K(SRC, true, 1, M, A, false, G)
/* the same four common-configuration kernels with the scattering queue (weakly scattering media, photon_kernel.cuh) */
#define MCXB_RDQ(SRC, M, A) MCXB_KQ(SRC, false, 0, M, A, false, false, MCXB_QUEUE_DEPTH), MCXB_KQ(SRC, true, 0, M, A, false, false, MCXB_QUEUE_DEPTH), \
                            MCXB_KQ(SRC, false, 1, M, A, false, false, MCXB_QUEUE_DEPTH), MCXB_KQ(SRC, true, 1, M, A, false, false, MCXB_QUEUE_DEPTH)
/* common-configuration kernels that read the record flags at run time (any -w / savedetflag) */
#define MCXB_D2(SRC, M, A) MCXB_K(SRC, false, 2, M, A, false, false), MCXB_K(SRC, true, 2, M, A, false, false)

static const KernelEntry entries[] = {
#if MCXB_INST_GROUP == 0
    MCXB_RD(srcPencil, uint8_t, double, false), MCXB_RD(srcPencil, uint8_t, float, false), MCXB_D2(srcPencil, uint8_t, double), MCXB_RDQ(srcPencil, uint8_t, double)
#elif MCXB_INST_GROUP == 1
    MCXB_RD(srcDisk, uint8_t, double, false), MCXB_RD(srcDisk, uint8_t, float, false), MCXB_D2(srcDisk, uint8_t, double), MCXB_RDQ(srcDisk, uint8_t, double)
#elif MCXB_INST_GROUP == 2
    MCXB_RD(srcPlanar, uint8_t, double, false), MCXB_RD(srcFourier, uint8_t, double, false), MCXB_RDQ(srcPlanar, uint8_t, double), MCXB_RDQ(srcFourier, uint8_t, double)
#elif MCXB_INST_GROUP == 3
    MCXB_RD(srcIsotropic, uint8_t, double, false), MCXB_RD(srcCone, uint8_t, double, false), MCXB_RDQ(srcIsotropic, uint8_t, double), MCXB_RDQ(srcCone, uint8_t, double)
#elif MCXB_INST_GROUP == 4
    MCXB_RD(srcAny, uint8_t, double, false), MCXB_RD(srcAny, uint8_t, float, false), MCXB_D2(srcAny, uint8_t, double), MCXB_D2(srcAny, uint8_t, float),
    MCXB_RDQ(srcAny, uint8_t, double)
#elif MCXB_INST_GROUP == 5
    MCXB_RD(srcAny, uint8_t, double, true), MCXB_RD(srcAny, uint8_t, float, true), MCXB_K(srcAny, true, 1, uint8_t, double, true, true)
#elif MCXB_INST_GROUP == 6
    MCXB_RD(srcAny, uint16_t, double, true), MCXB_RD(srcAny, uint16_t, float, true), MCXB_K(srcAny, true, 1, uint16_t, double, true, true)
#elif MCXB_INST_GROUP == 7
    MCXB_RD(srcAny, uint32_t, double, true), MCXB_RD(srcAny, uint32_t, float, true)
#elif MCXB_INST_GROUP == 8
    MCXB_RDX(uint8_t)
#elif MCXB_INST_GROUP == 9
    MCXB_RDX(uint16_t), MCXB_RDX(uint32_t)
#elif MCXB_INST_GROUP == 10
    MCXB_RDB(srcPencil, 0), MCXB_RDB(srcPencil, MCXB_QUEUE_DEPTH)
#elif MCXB_INST_GROUP == 11
    MCXB_RDB(srcAny, 0), MCXB_RDB(srcAny, MCXB_QUEUE_DEPTH)
#endif
};

} // namespace mcxb

#define MCXB_CAT2(a, b) a##b
#define MCXB_CAT(a, b) MCXB_CAT2(a, b)

extern "C" const mcxb::KernelEntry* MCXB_CAT(mcxb_kernel_group_, MCXB_INST_GROUP)(int* n) {
    *n = (int)(sizeof(mcxb::entries) / sizeof(mcxb::entries[0]));
    return mcxb::entries;
}
