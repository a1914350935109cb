/*
 * ref_variant.cpp -- one compile-time specialisation of the REFERENCE kernel source, built for the
 * host through clshim.h (TEST INFRASTRUCTURE ONLY; see oracle/README.md).
 *
 * The reference JIT-specialises mcx_main_loop with -D flags chosen per run
 * (reference src/mcx_host.cpp:857-971).  oracle/build_ref.py compiles this file once per
 * {source type} x {MCX_DO_REFLECTION} x {MCX_SAVE_DETECTORS} combination with the same flags and
 * a distinct REF_SUFFIX; the patched kernel text is generated into oracle/_ref/ at build time from
 * /root/reference/src/mcx_core.cl and is never committed.
 */
#include "clshim.h"

#ifndef REF_SUFFIX
    #error "REF_SUFFIX must be defined by the build script"
#endif
#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)

namespace REF_CAT(refk_, REF_SUFFIX) {
#include "mcx_core_patched.cl"
}

/* trajectory buffers of the calling host thread (set by ref_driver.cpp when MCX_DEBUG_MOVE is requested): passed this
 * way so that the kernel wrappers keep one signature */
extern thread_local unsigned int* mcxref_jumpdebug;
extern thread_local float* mcxref_debugdata;
extern thread_local const void* mcxref_smatrix;

extern "C" void REF_CAT(mcxref_kernel_, REF_SUFFIX)(
    const unsigned int* media, float* field, float* genergy, unsigned int* n_seed,
    float* n_det, const void* gproperty, float* srcpattern, const void* gdetpos,
    volatile unsigned int* gprogress, unsigned int* detectedphoton,
    unsigned long* gseeddata, float* ginvcdf, float* gangleinvcdf, void* sharedmem, const void* gcfg,
    float* replayweight, float* photontof, int* photondetid) {
    using namespace REF_CAT(refk_, REF_SUFFIX);
    mcx_main_loop(media, field, genergy, n_seed, n_det, (const float4*)gproperty, srcpattern,
                  (const float4*)gdetpos, gprogress, detectedphoton,
                  replayweight, photontof, photondetid,
                  (RandType*)gseeddata, mcxref_jumpdebug, mcxref_debugdata,
                  ginvcdf, gangleinvcdf, (RandType*)sharedmem, (float4*)mcxref_smatrix,
                  (const MCXParam*)gcfg);
}
