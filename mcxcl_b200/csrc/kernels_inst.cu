/*
 * kernels_inst.cu -- explicit instantiations of photon_kernel, one group per compilation
 * (nvcc ... -DMCXB_INST_GROUP=k).  See kernel_registry.h.
 *
 *   group 0: pencil beam,           8-bit media
 *   group 1: disk / ring source,    8-bit media
 *   group 2: planar + fourier,      8-bit media
 *   group 3: isotropic + cone,      8-bit media
 *   group 4: any source (run time), 8-bit media  (+ the instrumented variant that counts segments/deposits/scatters)
 *   group 5: any source (run time), 16-bit media (volumes with more than 127 labels)
 */
#include "kernel_registry.h"

#ifndef MCXB_INST_GROUP
    #error "compile with -DMCXB_INST_GROUP=<0..5>"
#endif

namespace mcxb {

#define MCXB_K(SRC, R, D, M, A, S) { SRC, R, D, sizeof(M) == 2, sizeof(A) == 8, S, photon_kernel<SRC, R, D, M, A, S>, #SRC "/" #R #D "/" #M "/" #A }
#define MCXB_RD(SRC, M, A) MCXB_K(SRC, false, false, M, A, false), MCXB_K(SRC, true, false, M, A, false), \
                           MCXB_K(SRC, false, true, M, A, false), MCXB_K(SRC, true, true, M, A, false)

static const KernelEntry entries[] = {
#if MCXB_INST_GROUP == 0
    MCXB_RD(srcPencil, uint8_t, double), MCXB_RD(srcPencil, uint8_t, float)
#elif MCXB_INST_GROUP == 1
    MCXB_RD(srcDisk, uint8_t, double), MCXB_RD(srcDisk, uint8_t, float)
#elif MCXB_INST_GROUP == 2
    MCXB_RD(srcPlanar, uint8_t, double), MCXB_RD(srcFourier, uint8_t, double)
#elif MCXB_INST_GROUP == 3
    MCXB_RD(srcIsotropic, uint8_t, double), MCXB_RD(srcCone, uint8_t, double)
#elif MCXB_INST_GROUP == 4
    MCXB_RD(srcAny, uint8_t, double), MCXB_RD(srcAny, uint8_t, float), MCXB_K(srcAny, true, true, uint8_t, double, true)
#elif MCXB_INST_GROUP == 5
    MCXB_RD(srcAny, uint16_t, double), MCXB_RD(srcAny, uint16_t, float), MCXB_K(srcAny, true, true, uint16_t, double, true)
#endif
};

} // namespace mcxb

#define MCXB_CAT2(a, b) a##b
#define MCXB_CAT(a, b) MCXB_CAT2(a, b)

extern "C" const mcxb::KernelEntry* MCXB_CAT(mcxb_kernel_group_, MCXB_INST_GROUP)(int* n) {
    *n = (int)(sizeof(mcxb::entries) / sizeof(mcxb::entries[0]));
    return mcxb::entries;
}
