/*
 * mcx_cuda_host.cpp -- the reference-side binding: exports the reference's own boundary symbols
 *
 *     void           mcx_run_simulation(Config* cfg, float* fluence, float* totalenergy);
 *     cl_platform_id mcx_list_gpu(Config* cfg, unsigned int* activedev, cl_device_id* activedevlist, GPUInfo** info);
 *     void           ocl_assess(int err, const char* file, const int linenum);
 *
 * (reference src/mcx_host.h:168-170) on top of the C ABI of include/mcxb200.h, so that the reference's
 * unchanged front-ends (src/mcxcl.c:35, src/pmcxcl.cpp:1147/1244/1609, src/mcxlabcl.cpp:139/264/345) and
 * its unchanged configuration / output code (src/mcx_utils.c) link against the B200 engine instead of
 * src/mcx_host.cpp + libOpenCL.  This file replaces src/mcx_host.cpp for the photon-transport path;
 * it is compiled against the reference's own headers where they lie (integration/build_cli.py), with
 * integration/clstub/CL/cl.h standing in for the OpenCL SDK header.
 *
 * Contract kept from the reference:
 *   inputs   everything is read from Config AFTER mcx_preprocess/mcx_validatecfg ran (the front-ends call
 *            them before reaching the boundary);
 *   outputs  cfg->exportfield (calloc'd if NULL, accumulated with +=, then normalised in place:
 *            src/mcx_host.cpp:1048-1054, 1292-1296, 1382-1465), cfg->exportdetected / seeddata (malloc /
 *            realloc: :1056-1058, 1218-1232), detectedcount, energytot/esc/abs, runtime, normalizer,
 *            his.*, maxgate; files through mcx_savedata / mcx_savedetphoton when parentid == mpStandalone
 *            (:1646-1664); the "simulated ... photon/ms" and "absorbed: ...%" report lines (:1677-1693);
 *   errors   mcx_error(id, msg, file, line): exit(id) standalone, exception under Python/MATLAB;
 *   devices  cfg->deviceid[] '1' flags select CUDA devices, cfg->workload[] weights split the photons
 *            (:650-662, 1011-1012); every device gets the next slice of ONE rand() stream (:759-768).
 *
 * Several selected GPUs (`-G 1101`, `-W a,b,c`) are driven from this one host process through
 * mcxb_run_simulation_multi: photon shards, one seed-stream slice per device, then an NCCL reduce of the volumes and
 * energies and a gather of the detected-photon records onto the first device, ONE read-back and one normalisation --
 * in place of the reference's per-device read-back and host-side summing (src/mcx_host.cpp:1218-1232, 1292-1306).
 * Only the `-D P` progress bar keeps the device-by-device form (it polls device 0 while the kernels run).
 */
#include "mcx_host.h"
#include "mcx_tictoc.h"
#include "mcx_const.h"
#include "mcxb200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>

namespace {

const char* const kEngineName = "B200-native CUDA engine (libmcxb200)";

void raise(int code, const char* file, int line) {
    const char* msg = mcxb_last_error();
    /* same convention as ocl_assess (src/mcx_host.cpp:213-217): the id handed to mcx_error is positive */
    mcx_error(code < 0 ? -code : code, (msg && msg[0]) ? msg : "CUDA engine error", file, line);
}

#define MCXB_TRY(call)                          \
    do {                                        \
        int rc__ = (call);                      \
        if (rc__ != MCXB_OK) {                  \
            raise(rc__, __FILE__, __LINE__);    \
        }                                       \
    } while (0)

mcxb_f4 f4(const float4& v) {
    mcxb_f4 r = {v.x, v.y, v.z, v.w};
    return r;
}

/* Config -> mcxb_config: field-by-field, no arithmetic (the reference's preprocessing already ran) */
void fill_config(const Config* cfg, mcxb_config* c) {
    memset(c, 0, sizeof(*c));
    c->abi_version = MCXB_ABI_VERSION;
    c->dimx = cfg->dim.x;
    c->dimy = cfg->dim.y;
    c->dimz = cfg->dim.z;
    c->vol = cfg->vol;
    c->unitinmm = cfg->unitinmm;
    c->mediaformat = cfg->mediabyte;       /* <= 4: label media; 99..104: the continuous formats packed by the front-end; 97: split-voxel media, two words per voxel */
    c->medianum = cfg->medianum;
    c->prop = reinterpret_cast<const mcxb_f4*>(cfg->prop);          /* Medium {mua,mus,g,n}, 16 bytes */
    c->srctype = cfg->srctype;
    c->src.pos = f4(cfg->srcpos);
    c->src.dir = f4(cfg->srcdir);
    c->src.param1 = f4(cfg->srcparam1);
    c->src.param2 = f4(cfg->srcparam2);
    c->extrasrclen = cfg->extrasrclen;
    c->srcdata = reinterpret_cast<const mcxb_source*>(cfg->srcdata);   /* ExtraSrc == 4 x float4 */
    c->srcid = cfg->srcid;
    c->srcnum = cfg->srcnum;
    c->srcpattern = cfg->srcpattern;

    if (cfg->srcpattern) {
        if (cfg->srctype == MCX_SRC_PATTERN3D) {
            c->srcpattern_len = (uint64_t)cfg->srcparam1.x * (uint64_t)cfg->srcparam1.y * (uint64_t)cfg->srcparam1.z * cfg->srcnum;
        } else {
            c->srcpattern_len = (uint64_t)cfg->srcparam1.w * (uint64_t)cfg->srcparam2.w * cfg->srcnum;
        }
    }

    c->nphase = cfg->nphase;
    c->nangle = cfg->nangle;
    c->invcdf = cfg->invcdf;
    c->angleinvcdf = cfg->angleinvcdf;
    c->detnum = cfg->detnum;
    c->detpos = reinterpret_cast<const mcxb_f4*>(cfg->detpos);
    c->issavedet = cfg->issavedet;
    c->savedetflag = cfg->savedetflag;
    c->maxdetphoton = cfg->maxdetphoton;
    c->issaveseed = cfg->issaveseed;
    c->issaveref = cfg->issaveref;
    c->tstart = cfg->tstart;
    c->tstep = cfg->tstep;
    c->tend = cfg->tend;
    c->nphoton = cfg->nphoton;
    c->seed = cfg->seed;
    c->isreflect = cfg->isreflect;
    memcpy(c->bc, cfg->bc, 12);
    c->isspecular = cfg->isspecular;
    c->minenergy = cfg->minenergy;
    c->gscatter = cfg->gscatter;
    c->maxvoidstep = cfg->maxvoidstep;
    c->voidtime = cfg->voidtime;
    c->outputtype = cfg->outputtype;
    c->isnormalized = 0;               /* normalised once, after every device has been summed */
    c->issave2pt = cfg->issave2pt;
    c->debuglevel = cfg->debuglevel & (MCX_DEBUG_RNG | MCX_DEBUG_MOVE | MCX_DEBUG_MOVE_ONLY);
    c->maxjumpdebug = cfg->maxjumpdebug;
    c->respin = cfg->respin;
    c->nthread = cfg->autopilot ? 0 : cfg->nthread;
    c->nblocksize = cfg->autopilot ? 0 : cfg->nblocksize;
    c->sched = MCXB_SCHED_DYNAMIC;
    c->accum = MCXB_ACCUM_F64;

    /* polarised light and RF: tables and scalars the front-end prepared (mcx_prep_polarized, src/mcx_utils.c:1483-1519) */
    c->polmedianum = cfg->smatrix ? cfg->polmedianum : 0;
    c->smatrix = reinterpret_cast<const mcxb_f4*>(cfg->smatrix);
    c->srciquv = f4(cfg->srciquv);
    c->omega = cfg->omega;

    if (cfg->seed == SEED_FROM_FILE) {
        /* photon replay: the records mcx_replayinit / mcx_replayprep prepared (src/mcx_utils.c:1355-1470) take the place
         * of the gseed / greplayw / greplaytof / greplaydetid buffers of src/mcx_host.cpp:722-737 */
        c->replay_seed = static_cast<const uint64_t*>(cfg->replay.seed);
        c->replay_weight = cfg->replay.weight;
        c->replay_tof = cfg->replay.tof;
        c->replay_detid = cfg->replay.detid;
        c->replaydet = cfg->replaydet;
    }
}

/* floats per detected-photon record (hostdetreclen, src/mcx_host.cpp:494-496) */
unsigned int record_length(const Config* cfg) {
    const unsigned int flag = cfg->issavedet ? (cfg->savedetflag & ((cfg->polmedianum && cfg->smatrix) ? 0xFFu : 0x7Fu)) : 0u;
    const unsigned int nmed = cfg->medianum - 1;
    return nmed * ((flag >> 1 & 1u) + (flag >> 2 & 1u) + (flag >> 3 & 1u)) + (flag & 1u) + 3 * ((flag >> 4 & 1u) + (flag >> 5 & 1u)) + (flag >> 6 & 1u) +
           4 * (flag >> 7 & 1u);
}

/* what this build's hot path does not cover is refused loudly, never approximated */
void check_supported(const Config* cfg) {
    if (cfg->mediabyte > 4 && (cfg->mediabyte < MEDIA_LABEL_HALF || cfg->mediabyte > MEDIA_AS_SHORT) && cfg->mediabyte != MEDIA_2LABEL_SPLIT) {
        mcx_error(-1, "mixed-label and two-word media formats are outside the photon-transport path of the CUDA engine", __FILE__, __LINE__);
    }

    if (cfg->seed == SEED_FROM_FILE && cfg->replay.seed == NULL) {
        mcx_error(-1, "replay needs the saved RNG states of the detected photons", __FILE__, __LINE__);
    }

    if (cfg->respin < 1) {
        mcx_error(-1, "negative respin is not supported by the CUDA engine", __FILE__, __LINE__);
    }

    if (cfg->istrajstokes && (cfg->debuglevel & (MCX_DEBUG_MOVE | MCX_DEBUG_MOVE_ONLY))) {
        mcx_error(-1, "trajectory records with Stokes vectors are outside the photon-transport path of the CUDA engine", __FILE__, __LINE__);
    }
}

}  // namespace

extern "C" void ocl_assess(int err, const char* file, const int linenum) {
    if (err != 0) {
        mcx_error(-err, "CUDA engine error", file, linenum);
    }
}

extern "C" cl_platform_id mcx_list_gpu(Config* cfg, unsigned int* activedev, cl_device_id* activedevlist, GPUInfo** info) {
    mcxb_gpuinfo dev[MAX_DEVICE];
    const int n = mcxb_list_gpu(dev, MAX_DEVICE);

    if (n < 0) {
        raise(n, __FILE__, __LINE__);
    }

    if (activedev) {
        *activedev = 0;
    }

    *info = (GPUInfo*)calloc(MAX_DEVICE, sizeof(GPUInfo));
    unsigned int active = 0;

    if (cfg->isgpuinfo && n > 0) {
        MCX_FPRINTF(stdout, S_YELLOW "Platform [0] Name %s\n" S_RESET, kEngineName);
    }

    for (int i = 0; i < n && i < MAX_DEVICE; i++) {
        GPUInfo g;
        memset(&g, 0, sizeof(g));
        strncpy(g.name, dev[i].name, sizeof(g.name) - 1);
        g.id = i + 1;
        g.devcount = n;
        g.platformid = 0;
        g.major = dev[i].major;
        g.minor = dev[i].minor;
        g.globalmem = dev[i].globalmem;
        g.constmem = dev[i].constmem;
        g.sharedmem = dev[i].sharedmem;
        g.regcount = dev[i].regcount;
        g.clock = dev[i].clock_khz / 1000;       /* the reference reports MHz (CL_DEVICE_MAX_CLOCK_FREQUENCY) */
        g.sm = dev[i].sm;
        g.core = dev[i].core;
        g.autoblock = dev[i].autoblock;
        g.autothread = dev[i].autothread;
        g.maxgate = cfg->maxgate;
        g.maxmpthread = dev[i].maxmpthread;
        g.iscpu = 0;
        g.vendor = dvNVIDIA;

        if (cfg->isgpuinfo) {
            MCX_FPRINTF(stdout, S_BLUE "============ GPU device ID %d [%d of %d]: %s  ============\n" S_RESET, i, i + 1, n, g.name);
            MCX_FPRINTF(stdout, " Device %d of %d:\t\t%s\n", i + 1, n, g.name);
            MCX_FPRINTF(stdout, " Compute units   :\t%d core(s)\n", g.sm);
            MCX_FPRINTF(stdout, " Global memory   :\t%.0f B\n", (double)g.globalmem);
            MCX_FPRINTF(stdout, " Local memory    :\t%.0f B\n", (double)g.sharedmem);
            MCX_FPRINTF(stdout, " Constant memory :\t%.0f B\n", (double)g.constmem);
            MCX_FPRINTF(stdout, " Clock speed     :\t%d MHz\n", g.clock);
            MCX_FPRINTF(stdout, " Compute Capacity:\t%d.%d\n", g.major, g.minor);
            MCX_FPRINTF(stdout, " Stream Processor:\t%d\n", g.core);
            MCX_FPRINTF(stdout, " Vendor name    :\t%s\n", "NVIDIA");
            MCX_FPRINTF(stdout, " Auto-thread    :\t%zu\n", g.autothread);
            MCX_FPRINTF(stdout, " Auto-block     :\t%zu\n", g.autoblock);
        }

        if (activedevlist != NULL) {
            if (cfg->deviceid[i] == '1') {
                memcpy((*info) + active, &g, sizeof(GPUInfo));
                activedevlist[active++] = (cl_device_id)(intptr_t)(i + 1);
            }
        } else {
            memcpy((*info) + active, &g, sizeof(GPUInfo));
            active++;
        }
    }

    if (activedev) {
        *activedev = active;
    }

    *info = (GPUInfo*)realloc(*info, std::max(1u, active) * sizeof(GPUInfo));

    if (cfg->isgpuinfo == 2 && cfg->parentid == mpStandalone) {
        exit(0);
    }

    return active ? (cl_platform_id)(intptr_t)1 : NULL;
}

extern "C" void mcx_run_simulation(Config* cfg, float* fluence, float* totalenergy) {
    (void)fluence;        /* never referenced by the reference either (callers pass NULL, src/mcxlabcl.cpp:345) */
    (void)totalenergy;

    cl_device_id devices[MAX_DEVICE];
    unsigned int workdev = 0;
    GPUInfo* gpu = NULL;
    mcx_list_gpu(cfg, &workdev, devices, &gpu);

    if (workdev == 0) {
        free(gpu);
        mcx_error(-99, "Specified GPU does not exist", __FILE__, __LINE__);
    }

    check_supported(cfg);

    cfg->maxgate = (unsigned int)((cfg->tend - cfg->tstart) / cfg->tstep + 0.5);
    const size_t dimxyz = (size_t)cfg->dim.x * cfg->dim.y * cfg->dim.z;
    const bool sharing = cfg->srctype == MCX_SRC_PATTERN && cfg->srcnum > 1;      /* photon sharing: one volume per pattern */
    const unsigned int nsrcvol = sharing ? cfg->srcnum : ((cfg->extrasrclen && cfg->srcid < 0) ? cfg->extrasrclen + 1 : 1);
    const bool replay = cfg->seed == SEED_FROM_FILE;
    const unsigned int nrepvol = (replay && cfg->replaydet == -1) ? std::max(1u, cfg->detnum) : 1u;     /* src/mcx_host.cpp:684-689 */
    /* RF outputs are complex, real volumes followed by imaginary volumes (src/mcx_host.cpp:1263-1276) */
    const bool isrfforward = cfg->omega > 0.f && !replay;
    const bool rfplanes = isrfforward || (replay && (cfg->outputtype == otRF || cfg->outputtype == otRFmus));
    const size_t planelen = dimxyz * cfg->maxgate * nsrcvol * nrepvol;
    const size_t fieldlen = planelen * (rfplanes ? 2 : 1);

    if (replay) {
        workdev = 1;        /* "replay should only work with a single device" (src/mcx_host.cpp:723) */
    }

    /* workload split (src/mcx_host.cpp:650-662, 1011-1012); the remainder goes to the first devices so
     * that exactly nphoton packets are launched */
    float fullload = 0.f;

    for (unsigned int i = 0; i < workdev; i++) {
        fullload += cfg->workload[i];
    }

    if (fullload < EPS) {
        for (unsigned int i = 0; i < workdev; i++) {
            cfg->workload[i] = (float)gpu[i].core;
            fullload += cfg->workload[i];
        }
    }

    std::vector<uint64_t> share(workdev);
    uint64_t assigned = 0;

    for (unsigned int i = 0; i < workdev; i++) {
        share[i] = (uint64_t)((double)cfg->nphoton * cfg->workload[i] / fullload);
        assigned += share[i];
    }

    for (unsigned int i = 0; assigned < cfg->nphoton; i = (i + 1) % workdev) {
        if (cfg->workload[i] > 0.f) {
            share[i]++;
            assigned++;
        }
    }

    std::vector<mcxb_sim*> sims;
    uint64_t seedskip = 0;
    int rc = MCXB_OK;
    const bool multi = workdev > 1 && !(cfg->debuglevel & (MCX_DEBUG_PROGRESS | MCX_DEBUG_RNG | MCX_DEBUG_MOVE | MCX_DEBUG_MOVE_ONLY));

    if (multi) {
        /* ---- all selected devices in ONE engine call: shards + NCCL exchange + one read-back ---- */
        const unsigned int reclen = record_length(cfg);

        if (cfg->exportfield == NULL && cfg->issave2pt) {
            cfg->exportfield = (float*)calloc(fieldlen, sizeof(float));
        }

        if (cfg->issavedet && cfg->exportdetected == NULL) {
            cfg->exportdetected = (float*)malloc(std::max<size_t>(1, (size_t)reclen * cfg->maxdetphoton) * sizeof(float));
        }

        if (cfg->issavedet && cfg->issaveseed && cfg->seeddata == NULL) {
            cfg->seeddata = malloc(std::max<size_t>(1, (size_t)cfg->maxdetphoton) * 16);
        }

        cfg->his.colcount = reclen;
        cfg->his.maxmedia = cfg->medianum - 1;
        cfg->his.detnum = cfg->detnum;
        cfg->his.srcnum = cfg->srcnum;
        cfg->his.savedetflag = cfg->savedetflag;
        cfg->his.totalsource = cfg->extrasrclen + 1;
        cfg->his.detected = 0;
        cfg->his.respin = cfg->respin;
        cfg->detectedcount = 0;
        cfg->energytot = cfg->energyesc = cfg->energyabs = 0.0;
        cfg->runtime = 0;

        mcx_printheader(cfg);
        MCX_FPRINTF(cfg->flog, "- code name: [%s] compiled for sm_100a\n", kEngineName);
        MCX_FPRINTF(cfg->flog, "- RNG: %s, photon scheduling: persistent threads with a per-GPU photon counter; %u devices combined over NVLink (peer memory; NCCL %d with MCXB_MULTI_EXCHANGE=nccl)\n",
                    MCX_RNG_NAME, workdev, mcxb_nccl_version());
        MCX_FPRINTF(cfg->flog, "initializing streams ...\t");
        mcx_flush(cfg);

        mcxb_config c;
        fill_config(cfg, &c);
        int devs[MAX_DEVICE];
        float wl[MAX_DEVICE];

        for (unsigned int i = 0; i < workdev; i++) {
            devs[i] = (int)(intptr_t)devices[i] - 1;
            wl[i] = cfg->workload[i];
        }

        mcxb_output out;
        mcxb_multi_info info;
        memset(&out, 0, sizeof(out));
        out.field = cfg->issave2pt ? cfg->exportfield : NULL;
        out.fieldlen = fieldlen;
        out.detphoton = cfg->issavedet ? cfg->exportdetected : NULL;
        out.seeddata = (cfg->issavedet && cfg->issaveseed) ? (uint64_t*)cfg->seeddata : NULL;
        const unsigned int tic0 = GetTimeMillis();
        MCX_FPRINTF(cfg->flog, "lauching mcx_main_loop for time window [%.1fns %.1fns] ...\n", cfg->tstart * 1e9, cfg->tend * 1e9);
        rc = mcxb_run_simulation_multi(&c, devs, (int)workdev, wl, &out, &info);

        if (rc == MCXB_OK) {
            for (unsigned int i = 0; i < workdev; i++) {
                MCX_FPRINTF(cfg->flog, "- [device %d(%d): %s] threadph=%d extra=%d np=%.1f nthread=%u nblock=%d repetition=%d\n",
                            i, gpu[i].id, gpu[i].name, (int)(info.share[i] / info.nthread[i]), (int)(info.share[i] % info.nthread[i]),
                            (double)info.share[i], info.nthread[i], (int)gpu[i].autoblock, 1);
            }

            cfg->runtime = std::max(1u, (unsigned int)(out.runtime_ms + 0.5f));
            MCX_FPRINTF(cfg->flog, "kernel complete:  \t%d ms\nretrieving flux ... \t", cfg->runtime);

            if (cfg->issavedet) {
                if (out.detected > cfg->maxdetphoton) {
                    MCX_FPRINTF(cfg->flog, S_RED "WARNING: the detected photon number is more than what your have specified (%u > %d), please use the -H option to specify a greater number\t" S_RESET,
                                out.detected, cfg->maxdetphoton);
                } else {
                    MCX_FPRINTF(cfg->flog, "detected " S_BOLD S_BLUE "%d photons" S_RESET ", total: " S_BOLD S_BLUE "%d" S_RESET "\t", out.detected, out.detected);
                }

                cfg->his.detected = out.detected;
                cfg->detectedcount = out.saved;
            }

            cfg->energytot = out.energytot;
            cfg->energyesc = out.energyesc;
            MCX_FPRINTF(cfg->flog, "transfer complete:        %d ms\n", GetTimeMillis() - tic0);
            mcx_flush(cfg);
        }
    } else {
        sims.assign(workdev, (mcxb_sim*)NULL);
    }

    /* ---- -D P / single device: one resident simulation per device; seed slices follow each other in ONE stream ---- */

    for (unsigned int i = 0; i < sims.size() && rc == MCXB_OK; i++) {
        mcxb_config c;
        fill_config(cfg, &c);
        c.nphoton = share[i];
        c.seed_skip = seedskip;
        rc = mcxb_sim_create(&c, (int)(intptr_t)devices[i] - 1, &sims[i]);

        if (rc == MCXB_OK) {
            seedskip += mcxb_sim_nthread(sims[i]);
            rc = mcxb_sim_reset(sims[i], NULL);
        }
    }

    if (rc == MCXB_OK && !multi) {
        if (cfg->exportfield == NULL && cfg->issave2pt) {
            cfg->exportfield = (float*)calloc(fieldlen, sizeof(float));
        }

        const unsigned int reclen = mcxb_sim_reclen(sims[0]);

        if (cfg->issavedet && cfg->exportdetected == NULL) {
            cfg->exportdetected = (float*)malloc(std::max<size_t>(1, (size_t)reclen * cfg->maxdetphoton) * sizeof(float));
        }

        if (cfg->issavedet && cfg->issaveseed && cfg->seeddata == NULL) {
            cfg->seeddata = malloc(std::max<size_t>(1, (size_t)cfg->maxdetphoton) * 16);
        }

        cfg->his.colcount = reclen;
        cfg->his.maxmedia = cfg->medianum - 1;
        cfg->his.detnum = cfg->detnum;
        cfg->his.srcnum = cfg->srcnum;
        cfg->his.savedetflag = cfg->savedetflag;
        cfg->his.totalsource = cfg->extrasrclen + 1;
        cfg->his.detected = 0;
        cfg->his.respin = cfg->respin;
        cfg->detectedcount = 0;
        cfg->energytot = cfg->energyesc = cfg->energyabs = 0.0;
        cfg->runtime = 0;

        mcx_printheader(cfg);
        MCX_FPRINTF(cfg->flog, "- code name: [%s] compiled for sm_100a\n", kEngineName);
        MCX_FPRINTF(cfg->flog, "- RNG: %s, photon scheduling: persistent threads with a per-GPU photon counter\n", MCX_RNG_NAME);
        MCX_FPRINTF(cfg->flog, "initializing streams ...\t");
        mcx_flush(cfg);

        /* ---- the reference's timing window: first enqueue ... all devices finished (:1078-1168) ---- */
        const unsigned int tic0 = GetTimeMillis();
        MCX_FPRINTF(cfg->flog, "lauching mcx_main_loop for time window [%.1fns %.1fns] ...\n", cfg->tstart * 1e9, cfg->tend * 1e9);

        /* `-r R`: the budget of every device in R batches, each with the next slice of the seed stream, accumulated on the
         * device and read back once (mcxb200.h, mcxb_config.respin) */
        const unsigned int respin = (unsigned int)std::max(1, cfg->respin);
        const uint64_t seedstride = seedskip;        /* threads of all devices: one slice per device and batch */
        float kernelms = 0.f;

        for (unsigned int i = 0; i < workdev; i++) {
            MCX_FPRINTF(cfg->flog, "- [device %d(%d): %s] threadph=%d extra=%d np=%.1f nthread=%u nblock=%d repetition=%d\n",
                        i, gpu[i].id, gpu[i].name, (int)(share[i] / respin / mcxb_sim_nthread(sims[i])), (int)(share[i] / respin % mcxb_sim_nthread(sims[i])),
                        (double)share[i], mcxb_sim_nthread(sims[i]), (int)gpu[i].autoblock, (int)respin);
        }

        for (unsigned int iter = 0; iter + 1 < respin && rc == MCXB_OK; iter++) {      /* all batches but the last */
            uint64_t skip = (uint64_t)iter * seedstride;
            float batchms = 0.f;
            MCX_FPRINTF(cfg->flog, "simulation run#%2d ... \n", iter + 1);

            for (unsigned int i = 0; i < workdev && rc == MCXB_OK; i++) {
                rc = mcxb_sim_set_photons(sims[i], share[i] / respin + (iter < share[i] % respin ? 1 : 0));

                if (rc == MCXB_OK && iter > 0) {
                    rc = mcxb_sim_reseed(sims[i], cfg->seed, skip);
                }

                if (rc == MCXB_OK) {
                    rc = mcxb_sim_launch(sims[i], NULL);
                }

                skip += mcxb_sim_nthread(sims[i]);
            }

            for (unsigned int i = 0; i < workdev && rc == MCXB_OK; i++) {
                batchms = std::max(batchms, mcxb_sim_last_kernel_ms(sims[i]));
            }

            kernelms += batchms;
        }

        {
            const unsigned int iter = respin - 1;
            uint64_t skip = (uint64_t)iter * seedstride;

            if (respin > 1) {
                MCX_FPRINTF(cfg->flog, "simulation run#%2d ... \n", iter + 1);
            }

            for (unsigned int i = 0; i < workdev && rc == MCXB_OK; i++) {
                if (respin > 1) {
                    rc = mcxb_sim_set_photons(sims[i], share[i] / respin + (iter < share[i] % respin ? 1 : 0));

                    if (rc == MCXB_OK) {
                        rc = mcxb_sim_reseed(sims[i], cfg->seed, skip);
                    }
                }

                if (rc == MCXB_OK) {
                    rc = mcxb_sim_launch(sims[i], NULL);
                }

                skip += mcxb_sim_nthread(sims[i]);
            }
        }

        if ((cfg->debuglevel & MCX_DEBUG_PROGRESS) && rc == MCXB_OK) {
            /* the -D P bar (src/mcx_host.cpp:1112-1141): the reference polls a mapped counter of device 0 every 100 ms;
             * here the photon counter of device 0 is copied over a side stream while the kernel runs */
            mcx_progressbar(-0.f, cfg);
            int finished = 0;

            while (rc == MCXB_OK && !finished) {
                uint64_t claimed = 0;
                rc = mcxb_sim_progress(sims[0], &claimed, &finished);

                if (rc == MCXB_OK && !finished) {
                    mcx_progressbar((float)((double)claimed / (double)std::max<uint64_t>(1, share[0] / respin)), cfg);
                    sleep_ms(50);
                }
            }

            mcx_progressbar(1.0f, cfg);
            MCX_FPRINTF(cfg->flog, "\n");
        }

        {
            float batchms = 0.f;

            for (unsigned int i = 0; i < workdev && rc == MCXB_OK; i++) {
                batchms = std::max(batchms, mcxb_sim_last_kernel_ms(sims[i]));     /* waits for device i */
            }

            kernelms += batchms;
        }

        const unsigned int toc = GetTimeMillis() - tic0;
        cfg->runtime = std::max(1u, std::max(toc, (unsigned int)(kernelms + 0.5f)));
        MCX_FPRINTF(cfg->flog, "kernel complete:  \t%d ms\nretrieving flux ... \t", toc);
        mcx_flush(cfg);

        /* ---- read back, sum over devices (:1172-1306) ---- */
        std::vector<float> detbuf;
        std::vector<uint64_t> seedbuf;

        for (unsigned int i = 0; i < workdev && rc == MCXB_OK; i++) {
            mcxb_output out;
            memset(&out, 0, sizeof(out));
            out.field = cfg->issave2pt ? cfg->exportfield : NULL;      /* accumulated with += by the engine */
            out.fieldlen = fieldlen;

            if (cfg->issavedet) {
                detbuf.resize(std::max<size_t>(1, (size_t)reclen * cfg->maxdetphoton));
                out.detphoton = detbuf.data();

                if (cfg->issaveseed) {
                    seedbuf.resize(std::max<size_t>(1, (size_t)cfg->maxdetphoton * 2));
                    out.seeddata = seedbuf.data();
                }
            }

            std::vector<float> trajbuf;

            if (cfg->debuglevel & (MCX_DEBUG_MOVE | MCX_DEBUG_MOVE_ONLY)) {
                trajbuf.resize(std::max<size_t>(1, (size_t)cfg->maxjumpdebug * MCXB_TRAJ_RECLEN));
                out.debugdata = trajbuf.data();
            }

            rc = mcxb_sim_fetch(sims[i], NULL, &out);

            if (rc == MCXB_OK && out.debugrecorded) {
                /* src/mcx_host.cpp:1173-1193 */
                if (out.debugrecorded > cfg->maxjumpdebug) {
                    MCX_FPRINTF(cfg->flog, S_RED "WARNING: the saved trajectory positions are more than what your have specified (%u > %d), please use the --maxjumpdebug option to specify a greater number\n" S_RESET,
                                out.debugrecorded, cfg->maxjumpdebug);
                } else {
                    MCX_FPRINTF(cfg->flog, "saved trajectory positions: %u, total: %d\t", out.debugrecorded, cfg->debugdatalen + out.debugrecorded);
                }

                cfg->exportdebugdata = (float*)realloc(cfg->exportdebugdata, (size_t)(cfg->debugdatalen + out.debugdatalen) * MCXB_TRAJ_RECLEN * sizeof(float));
                memcpy(cfg->exportdebugdata + (size_t)cfg->debugdatalen * MCXB_TRAJ_RECLEN, trajbuf.data(), (size_t)out.debugdatalen * MCXB_TRAJ_RECLEN * sizeof(float));
                cfg->debugdatalen += out.debugdatalen;
            }

            if (rc != MCXB_OK) {
                break;
            }

            if (cfg->issavedet) {
                if (out.detected > cfg->maxdetphoton) {
                    MCX_FPRINTF(cfg->flog, S_RED "WARNING: the detected photon number is more than what your have specified (%u > %d), please use the -H option to specify a greater number\t" S_RESET,
                                out.detected, cfg->maxdetphoton);
                } else {
                    MCX_FPRINTF(cfg->flog, "detected " S_BOLD S_BLUE "%d photons" S_RESET ", total: " S_BOLD S_BLUE "%d" S_RESET "\t", out.detected, cfg->detectedcount + out.detected);
                }

                cfg->his.detected += out.detected;

                if (cfg->exportdetected && out.saved) {
                    cfg->exportdetected = (float*)realloc(cfg->exportdetected, (size_t)(cfg->detectedcount + out.saved) * reclen * sizeof(float));
                    memcpy(cfg->exportdetected + (size_t)cfg->detectedcount * reclen, detbuf.data(), (size_t)out.saved * reclen * sizeof(float));

                    if (cfg->issaveseed && cfg->seeddata) {
                        cfg->seeddata = realloc(cfg->seeddata, (size_t)(cfg->detectedcount + out.saved) * 16);
                        memcpy((char*)cfg->seeddata + (size_t)cfg->detectedcount * 16, seedbuf.data(), (size_t)out.saved * 16);
                    }

                    cfg->detectedcount += out.saved;
                }
            }

            cfg->energytot += out.energytot;
            cfg->energyesc += out.energyesc;
        }

        MCX_FPRINTF(cfg->flog, "transfer complete:        %d ms\n", GetTimeMillis() - tic0);
        mcx_flush(cfg);
    }

    for (size_t i = 0; i < sims.size(); i++) {
        mcxb_sim_destroy(sims[i]);      /* full teardown before any error is raised (:35-49, 1846-1848) */
    }

    if (rc != MCXB_OK) {
        free(gpu);
        raise(rc, __FILE__, __LINE__);
        return;
    }

    std::vector<float> srcpw, srcetot, srceabs;

    /* ---- normalise once with the global launched energy (:1382-1465) ---- */
    if (cfg->issave2pt && cfg->isnormalized && !(cfg->debuglevel & MCX_DEBUG_RNG) && cfg->energytot > 0.0) {
        mcxb_config c;
        fill_config(cfg, &c);
        MCX_FPRINTF(cfg->flog, "normalizing raw data ...\t");
        cfg->energyabs += cfg->energytot - cfg->energyesc;
        const bool sens = cfg->outputtype == otJacobian || cfg->outputtype == otWP || cfg->outputtype == otDCS ||
                          cfg->outputtype == otWLTOF || cfg->outputtype == otWPTOF || cfg->outputtype == otRF || cfg->outputtype == otRFmus;

        if (replay && sens && cfg->replaydet == -1) {
            /* every detector at once: one scale per detector volume (src/mcx_host.cpp:1398-1421) */
            const size_t block = dimxyz * cfg->maxgate;

            for (int detid = 1; detid <= (int)cfg->detnum; detid++) {
                float scale = 0.f;

                for (size_t i = 0; i < cfg->nphoton; i++) {
                    if ((cfg->replay.detid[i] & 0xFFFF) == detid) {
                        scale += cfg->replay.weight[i];
                    }
                }

                if (scale > 0.f) {
                    scale = cfg->unitinmm / scale;
                }

                cfg->normalizer = scale;
                cfg->his.normalizer = scale;
                MCX_FPRINTF(cfg->flog, "normalization factor for detector %d alpha=%f\n", detid, scale);
                mcx_normalize(cfg->exportfield + (detid - 1) * block, scale, (int)block, cfg->isnormalized, 0, 1);

                if (rfplanes) {       /* src/mcx_host.cpp:1415-1417 */
                    mcx_normalize(cfg->exportfield + planelen + (detid - 1) * block, scale, (int)block, cfg->isnormalized, 0, 1);
                }
            }
        } else if (sharing) {
            /* per-pattern totals and scales, the reference's post-processing (src/mcx_host.cpp:1351-1380, 1436-1462) */
            const unsigned int psize = (unsigned int)((int)cfg->srcparam1.w * (int)cfg->srcparam2.w);
            const float ref = mcxb_normalizer(&c, cfg->energytot);
            srcpw.assign(cfg->srcnum, 0.f);
            srcetot.assign(cfg->srcnum, 0.f);
            srceabs.assign(cfg->srcnum, 0.f);

            for (unsigned int i = 0; i < cfg->srcnum; i++) {
                float kahanc = 0.f;

                for (unsigned int j = 0; j < psize; j++) {
                    mcx_kahanSum(&srcpw[i], &kahanc, cfg->srcpattern[j * cfg->srcnum + i]);
                }

                srcetot[i] = cfg->nphoton * srcpw[i] / (float)psize;
                kahanc = 0.f;

                if (cfg->outputtype == otEnergy) {
                    for (size_t j = 0; j < fieldlen / cfg->srcnum; j++) {
                        mcx_kahanSum(&srceabs[i], &kahanc, cfg->exportfield[j * cfg->srcnum + i]);
                    }
                } else {
                    for (unsigned int j = 0; j < cfg->maxgate; j++) {
                        for (size_t k = 0; k < dimxyz; k++) {
                            mcx_kahanSum(&srceabs[i], &kahanc, cfg->exportfield[(j * dimxyz + k) * cfg->srcnum + i] * mcx_updatemua((unsigned int)cfg->vol[k], cfg));
                        }
                    }
                }
            }

            for (unsigned int i = 0; i < cfg->srcnum; i++) {
                const float scale = psize / srcpw[i] * ref;

                if (i == 0) {
                    cfg->normalizer = scale;
                    cfg->his.normalizer = scale;
                }

                MCX_FPRINTF(cfg->flog, "source %d, normalization factor alpha=%f\n", i + 1, scale);
                mcx_normalize(cfg->exportfield, scale, (int)(fieldlen / cfg->srcnum), cfg->isnormalized, i, cfg->srcnum);
            }
        } else {
            const float scale = mcxb_normalizer(&c, cfg->energytot);
            cfg->normalizer = scale;
            cfg->his.normalizer = scale;
            MCX_FPRINTF(cfg->flog, "source 1, normalization factor alpha=%f\n", scale);
            mcx_normalize(cfg->exportfield, scale, (int)fieldlen, cfg->isnormalized, 0, 1);
        }
    } else {
        cfg->energyabs += cfg->energytot - cfg->energyesc;
    }

    /* ---- adjoint Jacobians: products of the normalised source and detector fluences (src/mcx_host.cpp:1468-1641); the
     *      per-voxel products run on the device (mcxb_adjoint_products), the scale factors are applied here ---- */
    if (cfg->issave2pt && MCX_IS_ADJOINT_TYPE(cfg->outputtype) && !replay && cfg->detdir != NULL && cfg->exportfield) {
        const unsigned int tic = GetTimeMillis();
        const int isdual = MCX_IS_DUAL_ADJOINT_TYPE(cfg->outputtype);
        const unsigned int Nd = cfg->detnum, Ns = cfg->extrasrclen + 1 - cfg->detnum;
        const size_t adjointlen = dimxyz * Ns * Nd, single = adjointlen * (isrfforward ? 2 : 1);
        const float Vvox = cfg->steps.x * cfg->steps.y * cfg->steps.z;
        const int dev0 = (int)(intptr_t)devices[0] - 1;
        std::vector<float> hmua, hsecond(single);
        const float* im = isrfforward ? cfg->exportfield + planelen : NULL;
        const bool wantprod = isdual || cfg->outputtype == otAdjoint;

        if (wantprod) {
            hmua.resize(single);
            rc = mcxb_adjoint_products(dev0, cfg->exportfield, im, cfg->dim.x, cfg->dim.y, cfg->dim.z, cfg->maxgate, Ns, Nd, 0, hmua.data());
        }

        if (rc == MCXB_OK && (isdual || cfg->outputtype != otAdjoint)) {
            rc = mcxb_adjoint_products(dev0, cfg->exportfield, im, cfg->dim.x, cfg->dim.y, cfg->dim.z, cfg->maxgate, Ns, Nd, 1, hsecond.data());
        }

        if (rc != MCXB_OK) {
            free(gpu);
            raise(rc, __FILE__, __LINE__);
            return;
        }

        /* 1 / (3 (1-g) mus^2) or 1 / (3 (1-g)^2 mus^2) per voxel (:1569-1583, 1612-1636) */
        auto opscale = [&](size_t vox, bool musp) -> float {
            const unsigned int medid = cfg->vol[vox] & 0xFF;

            if (medid < cfg->medianum) {
                const float mus = cfg->prop[medid].mus, onemg = 1.f - cfg->prop[medid].g;

                if (mus > 0.f && onemg > 0.f) {
                    return musp ? 1.f / (3.f * onemg * onemg * mus * mus) : 1.f / (3.f * onemg * mus * mus);
                }
            }

            return 0.f;
        };

        if (cfg->exportjacob) {
            free(cfg->exportjacob);
        }

        cfg->exportjacob = (float*)malloc(sizeof(float) * single * (isdual ? 2 : 1));

        if (isdual) {
            for (size_t k = 0; k < single; k++) {
                hmua[k] *= -Vvox;
                hsecond[k] *= -cfg->unitinmm;
            }

            if (cfg->outputtype == otAdjointMuaMusp) {
                for (size_t vox = 0; vox < dimxyz; vox++) {
                    const float f = opscale(vox, true);

                    for (unsigned int sd = 0; sd < Ns * Nd; sd++) {
                        hsecond[vox + (size_t)sd * dimxyz] *= f;

                        if (isrfforward) {
                            hsecond[vox + (size_t)sd * dimxyz + adjointlen] *= f;
                        }
                    }
                }
            }

            /* {mua Re, second Re, mua Im, second Im} (:1592-1600) */
            memcpy(cfg->exportjacob, hmua.data(), adjointlen * sizeof(float));
            memcpy(cfg->exportjacob + adjointlen, hsecond.data(), adjointlen * sizeof(float));

            if (isrfforward) {
                memcpy(cfg->exportjacob + 2 * adjointlen, hmua.data() + adjointlen, adjointlen * sizeof(float));
                memcpy(cfg->exportjacob + 3 * adjointlen, hsecond.data() + adjointlen, adjointlen * sizeof(float));
            }
        } else {
            const float* src = (cfg->outputtype == otAdjoint) ? hmua.data() : hsecond.data();
            const float adjscale = (cfg->outputtype == otAdjoint) ? -Vvox : -cfg->unitinmm;

            for (size_t k = 0; k < single; k++) {
                cfg->exportjacob[k] = src[k] * adjscale;
            }

            if (cfg->outputtype == otAdjointMus || cfg->outputtype == otAdjointMusp) {
                for (size_t vox = 0; vox < dimxyz; vox++) {
                    const float f = opscale(vox, cfg->outputtype == otAdjointMusp);

                    for (unsigned int sd = 0; sd < Ns * Nd; sd++) {
                        cfg->exportjacob[vox + (size_t)sd * dimxyz] *= f;

                        if (isrfforward) {
                            cfg->exportjacob[vox + (size_t)sd * dimxyz + adjointlen] *= f;
                        }
                    }
                }
            }
        }

        MCX_FPRINTF(cfg->flog, "adjoint Jacobian computation complete : %d ms\n", GetTimeMillis() - tic);
    }

#ifndef MCX_CONTAINER

    if (cfg->issave2pt && cfg->parentid == mpStandalone) {
        /* like the reference, the volume file of an RF run holds the real parts only (src/mcx_host.cpp:1646-1648) */
        MCX_FPRINTF(cfg->flog, "saving data to file ... %zu %d\t", planelen, cfg->maxgate);
        mcx_savedata(cfg->exportfield, planelen, cfg);
        MCX_FPRINTF(cfg->flog, "saving data complete\n\n");
        mcx_flush(cfg);
    }

    if (cfg->issavedet && cfg->parentid == mpStandalone && cfg->exportdetected) {
        cfg->his.unitinmm = cfg->unitinmm;
        cfg->his.savedphoton = cfg->detectedcount;
        cfg->his.totalphoton = cfg->nphoton;

        if (cfg->issaveseed) {
            cfg->his.seedbyte = 16;
        }

        cfg->his.detected = cfg->detectedcount;
        mcx_savedetphoton(cfg->exportdetected, cfg->seeddata, cfg->detectedcount, 0, cfg);
    }

    if ((cfg->debuglevel & (MCX_DEBUG_MOVE | MCX_DEBUG_MOVE_ONLY)) && cfg->parentid == mpStandalone && cfg->exportdebugdata) {
        /* src/mcx_host.cpp:1666-1672: <session>.mct, or the Trajectory block of <session>_detp.jdat */
        cfg->his.colcount = MCXB_TRAJ_RECLEN;
        cfg->his.savedphoton = cfg->debugdatalen;
        cfg->his.totalphoton = cfg->nphoton;
        cfg->his.detected = 0;
        mcx_savedetphoton(cfg->exportdebugdata, NULL, cfg->debugdatalen, 0, cfg);
    }

#endif

    MCX_FPRINTF(cfg->flog, "simulated %zu photons (%zu) with %d devices (repeat x%d)\nMCX simulation speed: " S_BOLD S_BLUE "%.2f photon/ms" S_RESET "\n",
                cfg->nphoton, cfg->nphoton, workdev, cfg->respin, (double)cfg->nphoton / std::max(1u, cfg->runtime));
    if (sharing && !srcetot.empty()) {
        for (unsigned int i = 0; i < cfg->srcnum; i++) {       /* src/mcx_host.cpp:1681-1688 */
            MCX_FPRINTF(cfg->flog, "source #%d total simulated energy: %.2f\tabsorbed: " S_BOLD S_BLUE "%5.5f%%" S_RESET "\n(loss due to initial specular reflection is excluded in the total)\n",
                        i + 1, srcetot[i], srceabs[i] / srcetot[i] * 100.f);
        }
    } else {
        MCX_FPRINTF(cfg->flog, "total simulated energy: %.2f\tabsorbed: " S_BOLD S_BLUE "%5.5f%%" S_RESET "\n(loss due to initial specular reflection is excluded in the total)\n",
                    cfg->energytot, (cfg->energytot - cfg->energyesc) / cfg->energytot * 100.f);
    }
    mcx_flush(cfg);
    free(gpu);
}
