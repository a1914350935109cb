/*
 * mcxb200.h -- C ABI of the B200-native photon-transport engine (libmcxb200.so).
 *
 * This is the drop-in boundary for ONE path of fangq/mcxcl: the work of the OpenCL kernel
 * mcx_main_loop (reference src/mcx_core.cl:2307-3307) together with the host runtime that feeds it
 * (reference src/mcx_host.cpp:438-1849, mcx_run_simulation) and device enumeration
 * (src/mcx_host.cpp:252-432, mcx_list_gpu).  Plain pointers and sizes only -- no OpenCL, torch or
 * C++ types appear in any signature.
 *
 * Every field of mcxb_config mirrors a member of the reference's `Config`
 * (src/mcx_utils.h:165-281) AFTER mcx_preprocess()/mcx_validatecfg() ran on it
 * (src/mcx_utils.c:1521-1809): source direction normalised, mua/mus scaled by unitinmm,
 * mus==0 -> 1e-10, detector voxels flagged in bit 31 of `vol` (mcx_maskdet, :4085-4198), boundary
 * conditions converted to integer codes, and -- for point-like sources -- the launch voxel index
 * and its label stored as raw uint bit patterns in srcparam2.z / srcparam2.w (:1718-1749).
 * INTEGRATION.md shows the 60-line adapter that fills this struct from a reference `Config*` and
 * thereby re-exports mcx_run_simulation / mcx_list_gpu / ocl_assess unchanged
 * (src/mcx_host.h:168-170).
 */
#ifndef MCXB200_H
#define MCXB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCXB_ABI_VERSION 4

/* error codes returned by every entry point (0 = success).  The reference reports OpenCL errors
 * negated through ocl_assess (src/mcx_host.cpp:213-217); CUDA runtime errors are reported the same
 * way: -(int)cudaError_t - 1000. */
#define MCXB_OK             0
#define MCXB_ERR_ARG       -1    /* invalid argument / unsupported configuration  */
#define MCXB_ERR_NODEVICE  -99   /* "Specified GPU does not exist" (src/mcx_host.cpp:533) */
#define MCXB_ERR_NOMEM     -6
#define MCXB_ERR_CUDA_BASE -1000

typedef struct mcxb_f4 {
    float x, y, z, w;
} mcxb_f4;

/* one source record == MCXSrc (src/mcx_host.h:99-104) == ExtraSrc (src/mcx_utils.h:101-106) */
typedef struct mcxb_source {
    mcxb_f4 pos;      /* xyz in voxel units (0-based), w = initial weight */
    mcxb_f4 dir;      /* unit vector, w = focal length / angular-mode selector */
    mcxb_f4 param1;
    mcxb_f4 param2;   /* .z/.w = launch voxel idx / label as uint bits for point-like sources */
} mcxb_source;

/* source types: same numbering as MCX_SRC_* (src/mcx_const.h:75-92) */
enum mcxb_srctype {
    MCXB_SRC_PENCIL = 0, MCXB_SRC_ISOTROPIC, MCXB_SRC_CONE, MCXB_SRC_GAUSSIAN, MCXB_SRC_PLANAR,
    MCXB_SRC_PATTERN, MCXB_SRC_FOURIER, MCXB_SRC_ARCSINE, MCXB_SRC_DISK, MCXB_SRC_FOURIERX,
    MCXB_SRC_FOURIERX2D, MCXB_SRC_ZGAUSSIAN, MCXB_SRC_LINE, MCXB_SRC_SLIT, MCXB_SRC_PENCILARRAY,
    MCXB_SRC_PATTERN3D, MCXB_SRC_HYPERBOLOID_GAUSSIAN, MCXB_SRC_RING
};

/* output types: same numbering as TOutputType (src/mcx_utils.h:58-60).  Flux / fluence / energy / otL are forward
 * outputs; Jacobian (absorption sensitivity), WP (scattering-count sensitivity), DCS (momentum transfer), their
 * time-of-flight weighted forms WLTOF / WPTOF and the RF Jacobians (RF: absorption, RFMUS: scattering, both complex,
 * omega > 0) exist only in replay mode (replay_seed != NULL).  The adjoint types are forward fluence runs over the
 * sources AND the detectors (the front-end appends the detectors as extra sources and sets srcid = -1,
 * src/mcx_utils.c:1902-1925) whose volumes mcxb_adjoint_products then multiplies pairwise. */
enum mcxb_outputtype { MCXB_OT_FLUX = 0, MCXB_OT_FLUENCE = 1, MCXB_OT_ENERGY = 2, MCXB_OT_JACOBIAN = 3, MCXB_OT_WP = 4, MCXB_OT_DCS = 5,
                       MCXB_OT_RF = 6, MCXB_OT_L = 7, MCXB_OT_RFMUS = 8, MCXB_OT_WLTOF = 9, MCXB_OT_WPTOF = 10,
                       MCXB_OT_ADJOINT = 11, MCXB_OT_ADJOINT_DCOEFF = 12, MCXB_OT_ADJOINT_MUS = 13, MCXB_OT_ADJOINT_MUSP = 14,
                       MCXB_OT_ADJOINT_MUA_D = 15, MCXB_OT_ADJOINT_MUA_MUSP = 16
                     };
#define MCXB_NANGLES 181             /* NANGLES: rows of one Mueller matrix table (src/mcx_const.h:67) */

/* boundary codes: TBoundary (src/mcx_utils.h:65) */
enum mcxb_mediaformat { MCXB_MEDIA_2LABEL_SPLIT = 97, MCXB_MEDIA_LABEL_HALF = 99, MCXB_MEDIA_AS_F2H = 100, MCXB_MEDIA_MUA_FLOAT = 101, MCXB_MEDIA_AS_HALF = 102,
                        MCXB_MEDIA_ASGN_BYTE = 103, MCXB_MEDIA_AS_SHORT = 104
                      };

enum mcxb_boundary { MCXB_BC_UNKNOWN = 0, MCXB_BC_REFLECT, MCXB_BC_ABSORB, MCXB_BC_MIRROR, MCXB_BC_CYCLIC };

/* photon -> thread scheduling */
enum mcxb_sched {
    MCXB_SCHED_DYNAMIC = 0,  /* persistent threads pull photons from a per-GPU counter (default) */
    MCXB_SCHED_STATIC  = 1   /* the reference's threadphoton/oddphoton split
                                (src/mcx_host.cpp:1011-1012, src/mcx_core.cl:2442) -- reproducible
                                stream->photon mapping, used by the oracle-parity tests */
};

/* fluence accumulator precision on the device.  The reference keeps fp32 accumulators accurate at 1e8+
 * photons by spilling +-1000 into a shadow copy of the volume using the value returned by the atomic
 * (MAX_ACCUM, src/mcx_core.cl:516, 2882-2887; folded by the host, src/mcx_host.cpp:1252-1258).  This
 * engine's default is fire-and-forget fp64 reductions (same memory as field+shadow, no return trip);
 * MCXB_ACCUM_F32 selects plain fp32 reductions (half the L2 footprint, fine below ~1e7 photons). */
enum mcxb_accum { MCXB_ACCUM_F64 = 0, MCXB_ACCUM_F32 = 1 };

#define MCXB_DEBUG_RNG       1u      /* MCX_DEBUG_RNG  (src/mcx_const.h:69) */
#define MCXB_DEBUG_MOVE      2u      /* MCX_DEBUG_MOVE: record photon trajectories (`-D M`, src/mcx_core.cl:929-959) */
#define MCXB_DEBUG_MOVE_ONLY 8u      /* MCX_DEBUG_MOVE_ONLY: trajectories only; the front-end turns volume and detector output off (src/mcx_utils.c:1552-1555) */
#define MCXB_DEBUG_STATS 0x10000u
#define MCXB_TRAJ_RECLEN 6           /* MCX_DEBUG_REC_LEN: {photon id (uint32 bits), x, y, z, weight, source id} */

typedef struct mcxb_config {
    uint32_t abi_version;          /* must be MCXB_ABI_VERSION */

    /* ---- domain: Config.dim / vol / unitinmm ---- */
    uint32_t dimx, dimy, dimz;
    const uint32_t* vol;           /* dimx*dimy*dimz words, x fastest; bits 0..30 label, bit 31 detector mask */
    float    unitinmm;
    /* Config.mediabyte (src/mcx_const.h:56-64): <= 4 = label media (the volume holds labels); otherwise one of the
     * continuous formats, where the 31 low bits of every word ENCODE the optical properties of the voxel, decoded per
     * segment like updateproperty (src/mcx_core.cl:1079-1193): 99 MEDIA_LABEL_HALF {half value, 2-bit slot, 14-bit
     * label}, 100 MEDIA_AS_F2H / 102 MEDIA_AS_HALF {half mua, half mus}, 101 MEDIA_MUA_FLOAT {float mua},
     * 103 MEDIA_ASGN_BYTE {mua, mus, g, n as bytes between prop[1] and prop[2]}, 104 MEDIA_AS_SHORT {mua, mus as
     * shorts between prop[1] and prop[2]}.
     * 97 MEDIA_2LABEL_SPLIT (split-voxel Monte Carlo, src/mcx_core.cl:1231-1344): vol holds TWO words per voxel, first
     * dimxyz words {lower label << 24 | upper label << 16 | px << 8 | py}, then dimxyz words {pz << 24 | nx << 16 | ny << 8 |
     * nz}, as mcx_preprocess leaves them (src/mcx_utils.c:1688-1712): a plane through (px, py, pz) / 255 inside the voxel
     * with normal (nx, ny, nz) * 2 / 255 - 1 separates the lower from the upper tissue; upper label 0 = ordinary voxel.
     * (96 two-word and 98 mixed-label media are formats no front-end of the reference produces / its kernel does not decode.) */
    uint32_t mediaformat;

    /* ---- media table: Config.prop / medianum ({mua,mus,g,n}, row 0 = background) ---- */
    uint32_t medianum;
    const mcxb_f4* prop;

    /* ---- sources: Config.srctype/srcpos/srcdir/srcparam1/srcparam2/srcdata/srcid/srcpattern ---- */
    int32_t  srctype;
    mcxb_source src;
    uint32_t extrasrclen;
    const mcxb_source* srcdata;    /* extrasrclen records or NULL */
    int32_t  srcid;                /* 0: pick randomly, one volume; k>0: only source k; -1: one volume per source */
    uint32_t srcnum;               /* number of patterns (photon sharing when >1; only 1 supported) */
    const float* srcpattern;       /* pattern / pattern3d intensity table or NULL */
    uint64_t srcpattern_len;       /* number of floats in srcpattern */
    /* optional inverse-CDF tables: Config.invcdf/nphase (scattering cos(theta)), Config.angleinvcdf/nangle
     * (launch zenith angle / pi); src/mcx_core.cl:2102-2118, 2475-2482 */
    uint32_t nphase, nangle;
    const float* invcdf;
    const float* angleinvcdf;

    /* ---- detectors: Config.detpos/detnum/issavedet/savedetflag/maxdetphoton ---- */
    uint32_t detnum;
    const mcxb_f4* detpos;         /* xyz centre (voxel units, 0-based), w = radius */
    int32_t  issavedet;
    uint32_t savedetflag;          /* bits D S P M X V W I (I = Stokes vector, polarised runs only), src/mcx_const.h:94-101 */
    uint32_t maxdetphoton;
    int32_t  issaveseed;
    int32_t  issaveref;

    /* ---- time gates ---- */
    float tstart, tstep, tend;

    /* ---- photon budget and RNG ---- */
    uint64_t nphoton;
    int32_t  seed;                 /* >0: srand(seed)-compatible stream; <=0: time(0) like the reference */
    uint64_t seed_skip;            /* number of per-thread seed records (4 x rand()) to discard first:
                                      rank r of a multi-GPU job passes r*nthread so that every GPU gets the
                                      next slice of ONE rand() stream (src/mcx_host.cpp:759-768) */

    /* ---- physics switches ---- */
    int32_t  isreflect;
    uint8_t  bc[12];               /* [0..5] boundary codes -x,-y,-z,+x,+y,+z; [6..11] detect-on-face flags */
    int32_t  isspecular;
    float    minenergy;
    uint32_t gscatter;
    int32_t  maxvoidstep;
    int32_t  voidtime;
    int32_t  outputtype;
    int32_t  isnormalized;
    int32_t  issave2pt;
    uint32_t debuglevel;           /* bit 0 (MCX_DEBUG_RNG): fill field with rand_uniform01 draws and return;
                                      MCXB_DEBUG_STATS: run the instrumented kernel that counts segments,
                                      deposits and scattering events (mcxb_output.stats) */

    /* ---- launch shape ---- */
    uint32_t nthread;              /* 0 = autopilot (persistent grid sized from the SM count) */
    uint32_t nblocksize;           /* 0 = autopilot */
    int32_t  sched;                /* enum mcxb_sched */
    int32_t  accum;                /* enum mcxb_accum: precision of the device fluence accumulators */

    /* ---- photon replay: Config.replay / replaydet with Config.seed == SEED_FROM_FILE (src/mcx_utils.h:133-142,
     *      src/mcx_host.cpp:722-737; kernel src/mcx_core.cl:1590-1596, 2567-2592, 2845-2858).  When replay_seed is
     *      set, photon i restarts its RNG stream from replay_seed[2i..2i+1] (the state mcxb_output.seeddata recorded
     *      for a detected photon) and nphoton is the number of records.
     *      Two defects of the reference's replay are NOT reproduced (tests/test_replay.py documents both): (i) its OpenCL
     *      kernel maps work-item t to record t*threadphoton + min(t, oddphoton-1) + k (src/mcx_core.cl:1591), which skips
     *      or repeats records whenever a work-item owns more than one photon -- here record i is photon i; (ii) at
     *      scattering sites it indexes replay_weight / replay_tof / replay_detid with f.w instead of f.w-1
     *      (src/mcx_core.cl:2569-2586 vs :2847), i.e. the WP / DCS / WPTOF deposits of photon i carry the weight of record
     *      i+1 -- here every output type uses record i. ---- */
    const uint64_t* replay_seed;   /* 2 words per photon, or NULL = forward simulation */
    const float*    replay_weight; /* detected weight of photon i (mcx_replayprep, src/mcx_utils.c:1355-1430) */
    const float*    replay_tof;    /* its time of flight in seconds: selects the time gate of the sensitivity outputs */
    const int32_t*  replay_detid;  /* its detector (low 16 bits, 1-based); needed when replaydet == -1 */
    int32_t         replaydet;     /* -1: one output volume per detector; otherwise one volume */

    /* ---- repetitions: Config.respin (`-r`).  nphoton is split into `respin` batches launched one after the other,
     *      each with the next slice of the seed stream (the reference reseeds from the continuing rand() stream,
     *      src/mcx_host.cpp:1319-1332); volumes, energies and detected photons accumulate over the batches and are
     *      read back and normalised ONCE.  0 and 1 both mean a single batch.  (The reference's own accumulation for
     *      respin > 1 adds the running device volume into the export buffer once per batch, :1280-1296, which
     *      over-counts by a factor 2R/(R+1) after normalisation; that arithmetic is not reproduced.) ---- */
    int32_t  respin;
    /* ---- trajectories: Config.maxjumpdebug, with MCXB_DEBUG_MOVE / MCXB_DEBUG_MOVE_ONLY in debuglevel.  One record per
     *      launch, per scattering event and per termination of every packet (src/mcx_core.cl:1497-1503, 2243-2249,
     *      2625-2632), up to maxjumpdebug records, in arbitrary order (one atomic counter) ---- */
    uint32_t maxjumpdebug;

    /* ---- polarised light: Config.polmedianum / smatrix / srciquv (src/mcx_utils.h:184-194).  smatrix holds, per
     *      non-background medium, MCXB_NANGLES rows {S11, S12, S33, S43} of its Mie scattering matrix on a uniform grid of
     *      the polar angle (mcx_prep_polarized, src/mcx_utils.c:1483-1519, run by the front-end; prop[] already carries the
     *      Mie mus and g).  With polmedianum > 0 every packet carries a Stokes vector (initially srciquv), scattering
     *      angles are drawn by rejection against the Mueller matrix (src/mcx_core.cl:2454-2468, 801-835) and savedetflag
     *      bit 7 (I) appends {I, Q, U, V} to the detected-photon record.  Label media, 3-D domains. ---- */
    uint32_t polmedianum;
    const mcxb_f4* smatrix;
    mcxb_f4  srciquv;
    /* ---- RF: Config.omega, the modulation angular frequency in rad/s.  In a forward run omega > 0 turns the packet
     *      weight into a complex number that rotates by omega*n/c0 per unit length (src/mcx_core.cl:2750-2763) and the
     *      output into TWO volume sets, real parts then imaginary parts (mcxb_output.fieldlen doubles; src/mcx_host.cpp:
     *      1263-1268); in a replay it drives the RF / RFMUS Jacobians, also two volume sets.  (The reference's kernel adds
     *      the imaginary part of RFMUS two volume sets behind the real one, :2601, past the end of the buffer its host
     *      allocates; here it goes into the second set, where the host looks for it.) ---- */
    float    omega;
} mcxb_config;

typedef struct mcxb_output {
    /* caller-owned buffers (may be NULL to skip that output) */
    float*    field;               /* fieldlen floats; results are ADDED to the existing contents, as the
                                      reference does with cfg->exportfield (src/mcx_host.cpp:1292-1296),
                                      then normalised in place when isnormalized */
    uint64_t  fieldlen;            /* in: capacity; out: dimxyz*maxgate*(number of output volumes) */
    float*    detphoton;           /* maxdetphoton*reclen floats */
    uint64_t* seeddata;            /* maxdetphoton*2 words when issaveseed */
    /* results */
    uint32_t  detected;            /* photons that hit a detector (may exceed maxdetphoton) */
    uint32_t  saved;               /* records actually stored = min(detected, maxdetphoton) */
    uint32_t  reclen;              /* floats per record (hostdetreclen, src/mcx_host.cpp:496) */
    uint32_t  maxgate;
    double    energytot, energyesc, energyabs;
    float     normalizer;
    float     runtime_ms;          /* kernel window only, the reference's `runtime` (src/mcx_host.cpp:1078-1168) */
    uint32_t  nthread, nblocksize; /* launch shape actually used */
    uint64_t  kernel_launches;     /* number of CUDA kernels this call launched */
    uint64_t  stats[3];            /* MCXB_DEBUG_STATS: ray segments, fluence deposits, scattering events */
    /* trajectories (caller-owned, maxjumpdebug * MCXB_TRAJ_RECLEN floats, or NULL); what the reference keeps in
     * cfg->exportdebugdata / debugdatalen (src/mcx_host.cpp:1173-1193) */
    float*    debugdata;
    uint32_t  debugrecorded;       /* positions the kernel wanted to record (may exceed maxjumpdebug) */
    uint32_t  debugdatalen;        /* records stored = min(debugrecorded, maxjumpdebug) */
} mcxb_output;

/* subset of GPUInfo (src/mcx_utils.h:143-163) */
typedef struct mcxb_gpuinfo {
    char     name[64];
    int32_t  id, devcount, major, minor;
    uint64_t globalmem, constmem, sharedmem;
    int32_t  regcount, clock_khz, sm, core;
    uint64_t autoblock, autothread;
    int32_t  maxmpthread;
    uint64_t l2cache;
} mcxb_gpuinfo;

/* ---- one-shot API: what mcx_run_simulation / mcx_list_gpu bind to ---------------------------- */

/* replaces mcx_list_gpu (src/mcx_host.cpp:252-432): fills up to `maxinfo` records, returns the
 * number of CUDA devices (>=0) or a negative error code. */
int mcxb_list_gpu(mcxb_gpuinfo* info, int maxinfo);

/* replaces mcx_run_simulation (src/mcx_host.cpp:438-1849) on CUDA device `device`: uploads the
 * HOST buffers of cfg, runs the photon kernel, reads back, folds and normalises into `out`. */
int mcxb_run_simulation(const mcxb_config* cfg, int device, mcxb_output* out);

/* ---- one call, several GPUs of one box, ONE host process ------------------------------------------------------
 * replaces the multi-device branch of mcx_run_simulation: the `-G 1101` device mask / `-W a,b,c` workload split
 * (src/mcx_host.cpp:650-662, 1011-1012, src/mcx_utils.c:4845-4865), one slice of the single rand() stream per device
 * (src/mcx_host.cpp:759-768), all devices inside one timing window (:1098-1168) -- and, where the reference reads every
 * device back and sums volumes, energies and detected-photon lists on the host (:1218-1232, 1292-1306), an exchange
 * over NVLink.  Default: peer memory -- ONE kernel on devices[0] reads the float32 volumes of all peers through
 * NVLink-mapped pointers and adds them into its own, the records (and RNG states) are copied peer to peer into the
 * tail of devices[0]'s buffer, the energy pairs and counts (a few bytes) go through the host; no communicator to build.
 * With MCXB_MULTI_EXCHANGE=nccl, or when a peer cannot be mapped: ncclReduce(float32 volume) + ncclReduce(float64
 * energy pair) to devices[0], ncclAllGather of the detected counts, grouped ncclSend / ncclRecv of the records.  Then
 * one device-to-host copy and one normalisation with the global launched energy.
 *   devices   CUDA ordinals, devices[0] collects the result;  workload: ndev weights or NULL (= equal shares)
 *   out       as for mcxb_run_simulation (runtime_ms = the slowest device's kernel window)
 *   info      optional per-device report
 * NCCL is bound at run time (dlopen libnccl.so.2); with ndev == 1 the call is mcxb_run_simulation. */
#define MCXB_MAX_DEVICES 16
typedef struct mcxb_multi_info {
    int32_t  ndev;
    int32_t  nccl_version;                   /* e.g. 22703; 0 = the exchange went over peer memory */
    uint64_t share[MCXB_MAX_DEVICES];        /* photons given to each device */
    uint32_t detected[MCXB_MAX_DEVICES];     /* photons each device detected */
    uint32_t nthread[MCXB_MAX_DEVICES];      /* RNG streams (= threads) of each device: its slice of the seed stream */
    float    kernel_ms[MCXB_MAX_DEVICES];
} mcxb_multi_info;
int mcxb_run_simulation_multi(const mcxb_config* cfg, const int* devices, int ndev, const float* workload, mcxb_output* out,
                              mcxb_multi_info* info);
/* the photon split used by the call above: nphoton * w_i / sum(w) rounded down, remainder to the first devices */
void mcxb_split_photons(uint64_t nphoton, const float* workload, int ndev, uint64_t* share);
/* NCCL version found at run time (0 = no usable libnccl.so.2) */
int mcxb_nccl_version(void);

/* last error message of the calling thread ("" if none) */
const char* mcxb_last_error(void);

/* Device and pinned buffers of finished simulations are kept per size and reused by the next call (front-ends
 * call mcx_run_simulation in loops; cudaFree / cudaFreeHost cost more than the rest of the host work).  This
 * returns every cached buffer to the driver, e.g. before handing the GPU to another library. */
void mcxb_release_cached_buffers(void);

/* ---- staged API: the same path with inputs resident in HBM (bench `value`, multi-GPU plumbing) */
typedef struct mcxb_sim mcxb_sim;

int  mcxb_sim_create(const mcxb_config* cfg, int device, mcxb_sim** sim);   /* H2D of media, tables, seeds */
int  mcxb_sim_reset(mcxb_sim* sim, void* cuda_stream);                      /* zero field / energy / counters, restore seeds */
int  mcxb_sim_launch(mcxb_sim* sim, void* cuda_stream);                     /* enqueue the photon kernel; asynchronous */
/* progress of a launch that may still be running: photons claimed so far (read over a side stream, does not wait
 * for the kernel) and whether the launch has finished -- what the reference's `-D P` bar polls through its mapped
 * gprogress word (src/mcx_host.cpp:1112-1141) */
int  mcxb_sim_progress(mcxb_sim* sim, uint64_t* claimed, int* finished);
int  mcxb_sim_set_photons(mcxb_sim* sim, uint64_t nphoton);                 /* change the photon budget of the next launch */
/* Config.respin on a resident simulation: nphoton in `respin` batches launched back to back WITHOUT a reset in between
 * (results accumulate on the device); batch k takes the seed-stream slice seed_skip + k * seed_stride (seed_stride = the
 * threads of all devices of the job).  Returns the summed kernel time.  A second mcxb_sim_launch without mcxb_sim_reset
 * accumulates in the same way. */
int  mcxb_sim_run_batches(mcxb_sim* sim, uint64_t nphoton, uint32_t respin, int32_t seed, uint64_t seed_skip, uint64_t seed_stride, float* kernel_ms);
int  mcxb_sim_reseed(mcxb_sim* sim, int32_t seed, uint64_t seed_skip);      /* new per-thread seed slice (rank r: skip r*nthread) */
int  mcxb_sim_finalize(mcxb_sim* sim, void* cuda_stream);                   /* accumulators -> float32 volume on the device; asynchronous */
int  mcxb_sim_fetch(mcxb_sim* sim, void* cuda_stream, mcxb_output* out);    /* finalize if needed, sync, D2H, add into out->field, normalise */
/* raw device pointers, valid after mcxb_sim_finalize, so that a host framework (torch.distributed / NCCL)
 * can reduce the volume and the energy totals across GPUs in place before rank 0 calls mcxb_sim_fetch
 * (the reference sums per-device results on the host, src/mcx_host.cpp:1292-1306) */
void*    mcxb_sim_field_devptr(mcxb_sim* sim);      /* float32[fieldlen] raw (un-normalised) deposits */
void*    mcxb_sim_energy_devptr(mcxb_sim* sim);     /* double[2] = {escaped, launched} */
void*    mcxb_sim_detphoton_devptr(mcxb_sim* sim);  /* float32[maxdetphoton*reclen] */
void*    mcxb_sim_detcount_devptr(mcxb_sim* sim);   /* uint32[1] */
void*    mcxb_sim_seeddata_devptr(mcxb_sim* sim);   /* uint64[maxdetphoton*2] or NULL */
uint64_t mcxb_sim_fieldlen(mcxb_sim* sim);
uint32_t mcxb_sim_reclen(mcxb_sim* sim);
uint32_t mcxb_sim_nthread(mcxb_sim* sim);           /* threads (= RNG streams) this simulation uses */
uint32_t mcxb_sim_acc_copies(mcxb_sim* sim);        /* replicated accumulator volumes on the device (summed by finalize) */
const char* mcxb_sim_kernel_name(mcxb_sim* sim);    /* which specialisation was selected */
float mcxb_sim_last_kernel_ms(mcxb_sim* sim);       /* CUDA-event time of the most recent launch (syncs on it) */
void mcxb_sim_destroy(mcxb_sim* sim);

/* the post-kernels of the adjoint output types (mcx_adjoint_kernel / mcx_adjoint_dcoeff_kernel, src/mcx_core.cl:3393-3512,
 * launched by src/mcx_host.cpp:1498-1537) on CUDA device `device`.  field_re (and field_im for an RF run, else NULL)
 * hold the NORMALISED fluence of ns source volumes followed by nd detector volumes, maxgate gates each (host memory,
 * dimxyz * maxgate * (ns + nd) floats).  out receives dimxyz * ns * nd floats, pair (s, d) at (s * nd + d) * dimxyz --
 * gates summed, then phi_src * phi_det (gradient == 0) or grad phi_src . grad phi_det by second-order finite differences
 * in voxel units (gradient != 0) -- followed, when field_im is given, by the same number of imaginary parts.  The scale
 * factors (-Vvox, -unitinmm, 1 / (3 (1-g) mus^2) ...) stay with the caller (src/mcx_host.cpp:1560-1640). */
int mcxb_adjoint_products(int device, const float* field_re, const float* field_im, uint32_t dimx, uint32_t dimy, uint32_t dimz,
                          uint32_t maxgate, uint32_t ns, uint32_t nd, int gradient, float* out);

/* host-side normalisation shared by fetch and by the multi-GPU reducer (src/mcx_host.cpp:1382-1465):
 * returns the scale factor for the given totals */
float mcxb_normalizer(const mcxb_config* cfg, double energytot);

/* ---- unit-level hooks used by the parity tests (they run the SAME device functions the photon
 *      kernel is built from) ------------------------------------------------------------------ */

/* n RNG streams seeded with seeds[4*i..4*i+3] (xorshift128p_seed, src/mcx_core.cl:709-712); writes
 * ndraw floats per stream to out[i*ndraw + k] and the final 128-bit state to state_out[2*i..] */
int mcxb_test_rng(int device, const uint32_t* seeds, uint32_t n, uint32_t ndraw, float* out, uint64_t* state_out);

/* voxel traversal of n fixed rays for nstep segments each, no scattering: per step writes
 * dist, p.xyz (as float bit patterns), voxel ijk, face id and idx1d.  musp = the mus' used in the
 * (dist*mus')/mus' round trip of src/mcx_core.cl:2678-2680. */
typedef struct mcxb_trace_step {
    float    dist;
    float    px, py, pz;
    int16_t  ix, iy, iz, face;
    uint32_t idx1d;
} mcxb_trace_step;
int mcxb_test_trace(int device, const mcxb_f4* p0, const mcxb_f4* v0, uint32_t n, uint32_t nstep,
                    uint32_t dimx, uint32_t dimy, uint32_t dimz, float musp, mcxb_trace_step* out);

/* scalar helpers: mcx_nextafterf (src/mcx_core.cl:965-973), reflectcoeff (:1057-1075) */
int mcxb_test_scalar(int device, const float* a, const int32_t* dir, uint32_t n, float* nextafter_out,
                     const mcxb_f4* v, const float* n1, const float* n2, const int32_t* face, uint32_t m, float* rcoef_out);

/* rotatevector (src/mcx_core.cl:1025-1042) and transmit (:1044-1055) on n vectors, in place (fast tier:
 * compared with a tolerance, not bit-exactly) */
int mcxb_test_rotate(int device, mcxb_f4* v, const float* stheta, const float* ctheta, const float* sphi, const float* cphi, uint32_t n);
int mcxb_test_refract(int device, mcxb_f4* v, const float* n1, const float* n2, const int32_t* face, uint32_t n);

/* L2 reduction microbenchmark (the measured denominator of the RED roofline in bench.py): nblock*256
 * threads each issue `iters` red.global.add of 4- or 8-byte elements to pseudo-random addresses inside a
 * span_elems buffer; hot_permille of them are clustered on 64 elements.  Returns the mean time of one
 * launch in ms and the number of reductions per launch. */
int mcxb_bench_red(int device, int elem_bytes, uint64_t span_elems, uint32_t nblock, uint32_t iters,
                   uint32_t hot_permille, uint32_t repeats, float* ms_out, uint64_t* ops_out);

/* measurement hook (profiles/r2_deposit_variants.md): only the SMs selected by `mode` (0 all, 1/2 lower/upper half of the
 * SM ids, 3/4 even/odd ids) issue fp64 reductions into one 32-byte sector; run under ncu to see where the L2 "lookup miss"
 * of a reduction comes from.  Returns the reductions issued and the sum that arrived (they must be equal). */
int mcxb_bench_red_die(int device, int mode, uint32_t nblock, uint32_t iters, uint64_t* issued_out, double* sum_out);

/* glibc rand()-compatible seed table used by mcxb_sim_create (src/mcx_host.cpp:696-700, 759-768) */
void mcxb_fill_seeds(int32_t seed, uint64_t skip_records, uint64_t nrecords, uint32_t* out4);

#ifdef __cplusplus
}
#endif
#endif
