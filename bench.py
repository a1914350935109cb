#!/usr/bin/env python3
"""bench.py -- photons/ms of the photon-transport hot path on B200s (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload cube60b] [--photons P]

A "step" is one complete simulation of P photons per GPU of the workload (default: cube60b, 1e8 photons,
the configuration BASELINE.json quotes the metric on).  Printed by rank 0 as ONE JSON line:

  value          whole-job photons/ms with the inputs (media volume, tables, seeds) resident in HBM:
                 reset + photon kernel + accumulator finalisation (+ the NCCL combine step when N>1),
                 timed with CUDA events, L2 flushed between steps, max over ranks
  e2e            the same metric through the public host-buffer call (mcxb_run_simulation via
                 mcxcl_b200.engine / multigpu.run_distributed): H2D of volume+tables+seeds, kernel, D2H of the
                 fluence volume and detected photons, normalisation -- all inside the timed region
  roofline       the BINDING ceiling of the photon kernel on top: SM issue slots (warp instructions per photon from
                 the committed ncu capture x photons / kernel time, against 148 SMs x 4 schedulers x the sampled SM
                 clock), with lane occupancy and the thread-level fraction; beside it, as scalars and sub-objects, the
                 contract's HBM figure (algorithmic bytes / kernel time vs the measured copy bandwidth -- NOT binding:
                 the working set is L2-resident) and L2 reduction throughput (RED microbenchmark run live)
  cpu_baseline   the reference kernel source built for the host (oracle/_ref), all host cores, bounded sample
  parity_check   every step checks itself: energytot == N x nphoton exactly, whole-job detected count, records gathered
                 == sum of the per-rank counts, absorbed fraction within 0.5 % of the committed reference series
  extra          the same measurement (value, e2e, parity_check) for colin27, for digimouse (the deck the reference ships: a
                 fourier wide-field source, one gate) and for digimouse_tg (this repository's reading of BASELINE.json's
                 "multi-source time-gated": 4 pencil sources, one volume each, 10 gates = 392 M accumulators, DESIGN.md
                 section 7) at 1.25e8 photons per GPU: at 8 GPUs these are the 1e9-photon runs of BASELINE.json's configs,
                 colin27 being its scaling target

--impl reference times that CPU implementation alone (rank 0 only under torchrun).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

PUBLISHED_PHOTONS_PER_MS = 12953.37     # reference README.md:618 (cube60b-like deck, 1e7 photons, Titan V class, OpenCL)
# work per photon probed with the reference kernel source
WORK = {"cube60": dict(seg=196.7, dep=117.8, sca=79.7), "cube60b": dict(seg=322.7, dep=193.4, sca=130.0),
        "skinvessel": dict(seg=343.9, dep=307.2, sca=35.7), "colin27": dict(seg=1880.5, dep=233.4, sca=1647.4)}
FALLBACK_HBM_GBS = 6650.0
N_SM, SCHEDULERS_PER_SM = 148, 4


def kernel_counters(workload):
    """per-photon instruction counts and DRAM traffic of the photon kernel from the committed ncu capture
    (profiles/kernel_counters.json, written by tools/ncu_counters.py from the .ncu-rep of the same kernel version)"""
    path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    try:
        return json.load(open(path)).get(workload)
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm_GBs"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons DURING the timed region: ONE `nvidia-smi -lms 200` process, as in the
    profiling recipe (B200_PROFILING.md).  Spawning a new nvidia-smi per sample re-initialises NVML every time and was
    measured to slow the photon kernel by 3.5 % (362 ms vs 348 ms for cube60b 1e8)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
        super().start()

    def run(self):
        if self.proc is None:
            return
        for line in self.proc.stdout:
            line = line.strip()
            if line:
                self.rows.append([x.strip() for x in line.split(",")])

    def summary(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        self.join(timeout=6)
        if not self.rows:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        pw = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                    power_w=max(pw) if pw else None, samples=len(self.rows), reasons=reasons)


def host_threads():
    """threads the CPU arm may use: the cores this process is allowed on, NOT what OpenMP inherits
    (torch.distributed.run exports OMP_NUM_THREADS=1 to every rank when it starts more than one)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def bounded_cpu_run(workload, photons, hostthreads=0):
    """the reference kernel source on the host cores (oracle/_ref), or the C restatement when it is absent"""
    from mcxcl_b200 import benchmarks, hostcfg
    from oracle import loader
    kind = "reference" if loader.have_ref() else "port"
    chk = loader.ref() if kind == "reference" else loader.port()
    p = hostcfg.prepare(benchmarks.get(workload, photons))
    threads = hostthreads if hostthreads > 0 else host_threads()
    work = 64 * threads * 4
    t0 = time.perf_counter()
    o = chk.run(p, work, hostthreads=threads)          # explicit count: `#pragma omp parallel num_threads(threads)`
    wall = (time.perf_counter() - t0) * 1e3
    return dict(kind=kind, cores=threads, photons=int(photons), kernel_ms=o["runtime_ms"], wall_ms=wall, work_items=work,
                value=photons / o["runtime_ms"], absorbed=o["absorbed"])


def run_reference_arm(args, rank):
    """the reference's own CPU implementation of the path, all host cores, rank 0 only (the CPU arm does not use the
    GPUs: its value is the same at every N)"""
    if rank != 0:
        return
    photons = int(args.ref_photons)
    vals = []
    info = None
    for i in range(args.warmup + args.steps):
        info = bounded_cpu_run(args.workload, photons)
        if i >= args.warmup:
            vals.append(info["kernel_ms"])
    ms = float(np.mean(vals))
    v = photons / ms
    line = dict(impl="reference", metric="photons/ms", value=v, unit="photons/ms", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=dict(workload=args.workload, nphoton_per_step=photons, threads=info["cores"],
                            omp_num_threads_env=os.environ.get("OMP_NUM_THREADS"),
                            note="bounded sample of the workload; CPU implementation = the reference's kernel source (src/mcx_core.cl) compiled for the host through oracle/clshim.h, OpenMP with an explicit thread count = the cores this process may run on"),
                cpu_baseline=dict(value=v, unit="photons/ms", cores=info["cores"], kind=info["kind"], sample="%d photons of %s per step, %d work-items" % (photons, args.workload, info["work_items"])),
                e2e=dict(value=v, unit="photons/ms", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


def golden_absorbed(workload):
    """mean absorbed fraction of the committed reference series (tests/golden/ref_stats_<deck>.npz; a fixture, not the oracle)"""
    path = os.path.join(ROOT, "tests", "golden", "ref_stats_%s.npz" % workload)
    try:
        return float(np.load(path)["absorbed"].mean())
    except Exception:
        return None


class Runner:
    """one workload on this rank's GPU: device-resident steps (`value`) and host-buffer calls (`e2e`)"""

    def __init__(self, ctx, workload, nph):
        import torch
        from mcxcl_b200 import benchmarks, engine, hostcfg
        self.ctx, self.workload, self.nph = ctx, workload, int(nph)
        cfg = benchmarks.get(workload, self.nph)
        if cfg.get("detpos") is not None and cfg.get("issavedet", 1):
            # the reference only warns when the detected-photon buffer overflows (src/mcx_host.cpp:1207-1210) and tells the
            # user to raise -H: sized here so that rank 0 can hold every rank's records (checked in parity_check)
            cfg["maxdetphoton"] = int(max(1000000, 0.006 * self.nph * ctx["world"] + 100000))
        self.p = hostcfg.prepare(cfg)
        self.engine, self.torch = engine, torch

    def device_steps(self, warmup, steps, sampler=None):
        ctx, torch = self.ctx, self.torch
        from mcxcl_b200 import multigpu
        dist, rank, world, device = ctx["dist"], ctx["rank"], ctx["world"], ctx["device"]
        sim = self.engine.Simulation(self.p, ctx["local"])
        skip = multigpu.seed_offsets(dist, world, sim.nthread, device)[rank] if world > 1 else 0
        sim.reseed(self.p.c.seed, skip)
        flush = ctx["flush"]()
        kernel_ms, step_ms = [], []
        launches = 0
        gathered = None

        def step(timed):
            nonlocal launches, gathered
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sim.reset()
            sim.launch()
            if world > 1:
                gathered = multigpu.combine(sim, dist, rank, world, device)
            else:
                sim.finalize()
            e1.record()
            e1.synchronize()
            if timed:
                step_ms.append(e0.elapsed_time(e1))
                kernel_ms.append(sim.kernel_ms())
                launches += 2

        for _ in range(warmup):
            step(False)
        ctx["barrier"]()
        if sampler is not None:
            sampler.rows.clear()                   # keep only what is sampled during the timed region
        for _ in range(steps):
            step(True)
        ctx["barrier"]()
        clocks = sampler.summary() if sampler is not None else None
        total_ms = ctx["maxreduce"](sum(step_ms))
        kern_ms = ctx["maxreduce"](float(np.mean(kernel_ms)))
        res = sim.fetch() if rank == 0 else None
        if rank == 0 and gathered is not None and gathered[0] is not None:
            res["detected"] = int(sum(gathered[2]))
            res["saved"] = int(gathered[0].numel() // max(1, sim.reclen))
        out = dict(total_ms=total_ms, kernel_ms=kern_ms, kernel_ms_steps=[round(x, 2) for x in kernel_ms], launches=launches, clocks=clocks,
                   res=res, nthread=sim.nthread, kname=sim.kernel_name, copies=sim.acc_copies, reclen=sim.reclen)
        sim.close()
        return out

    def e2e_steps(self, warmup, steps):
        ctx = self.ctx
        from mcxcl_b200 import multigpu
        world = ctx["world"]
        p_all = self.p.clone_for(nphoton=self.nph * world)      # the whole job's budget; every rank takes its share of it
        ms, out = [], None
        for i in range(warmup + steps):
            ctx["barrier"]()
            t0 = time.perf_counter()
            if world > 1:
                out = multigpu.run_distributed(p_all)
            else:
                out = self.engine.run_prepared(self.p, ctx["local"])
            ctx["barrier"]()
            if i >= warmup:
                ms.append((time.perf_counter() - t0) * 1e3)
        return ctx["maxreduce"](sum(ms)), out

    def parity_check(self, dev, e2e_out):
        """self-check of one bench step against exact invariants and the committed reference series (rank 0)"""
        world, nph, p = self.ctx["world"], self.nph, self.p
        res, chk = dev["res"], {}
        unit_weight = p.c.srctype == 0 and p.c.extrasrclen == 0
        chk["energytot"] = res["energytot"]
        if unit_weight:
            chk["energytot_exact"] = bool(res["energytot"] == float(nph) * world)     # N x nphoton unit-weight packets, summed over ranks
        gold = golden_absorbed(self.workload)
        chk["absorbed"], chk["golden_absorbed"] = res["absorbed"], gold
        if gold:
            chk["absorbed_rel_err"] = abs(res["absorbed"] - gold) / gold
            chk["absorbed_within_0.5pct"] = bool(chk["absorbed_rel_err"] < 0.005)
        if p.c.issavedet and p.c.detnum:
            chk["detected_whole_job"] = int(res["detected"])
            chk["records_gathered"] = int(res["saved"])
            chk["records_ok"] = bool(res["saved"] == min(res["detected"], p.c.maxdetphoton))
            if e2e_out is not None and e2e_out.get("detp") is not None:
                d = e2e_out["detp"]
                chk["e2e_records"] = int(d.shape[0])
                chk["e2e_records_ok"] = bool(d.shape[0] == min(e2e_out["detected"], p.c.maxdetphoton))
                if p.c.savedetflag & 1:
                    chk["e2e_detector_ids_ok"] = bool(set(np.unique(d[:, 0].astype(np.int64) & 0xFFFF)) <= set(range(1, p.c.detnum + 1)))
                rate = res["detected"] / (float(nph) * world)
                chk["detected_per_photon"] = rate
                chk["e2e_detected_rate_ok"] = bool(abs(e2e_out["detected"] / (float(nph) * world) - rate) < 6 * np.sqrt(rate / (nph * world)) + 1e-9)
        if e2e_out is not None:
            chk["e2e_absorbed"] = e2e_out["absorbed"]
            if unit_weight:
                chk["e2e_energytot_exact"] = bool(e2e_out["energytot"] == float(nph) * world)
        chk["ok"] = all(v for k, v in chk.items() if isinstance(v, bool))
        return chk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cube60b")
    ap.add_argument("--photons", type=float, default=1e8, help="photons per GPU per step")
    ap.add_argument("--ref-photons", type=float, default=1e6, help="photons per step of the CPU reference arm")
    ap.add_argument("--cpu-photons", type=float, default=2e6, help="bounded sample of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--extra", default="colin27,digimouse,digimouse_tg", help="comma-separated extra workloads reported under `extra` (\"\" = none)")
    ap.add_argument("--extra-photons", type=float, default=1.25e8, help="photons per GPU per step of the extra workloads (1.25e8 x 8 GPUs = the 1e9-photon colin27 run)")
    ap.add_argument("--extra-steps", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "native" and args.warmup < 3:
        sys.stderr.write("bench.py: --warmup %d raised to 3 (timing rules: at least 3 warm-up steps)\n" % args.warmup)
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from mcxcl_b200 import abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the photon-transport engine has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = abi.load()
    nph = int(args.photons)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    flushbuf = []

    def flush():
        # the L2-flush buffer is allocated AFTER the simulation's buffers: with the 384 MB block allocated first the driver
        # places the (2 MB) working set differently and the same kernel was measured 3.5 % slower (360.4 ms vs 348.0 ms,
        # profiles/README.md "placement"); a front-end that calls the engine has no such block in the way
        if not flushbuf:
            flushbuf.append(torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=device))      # > 126 MB of L2
        return flushbuf[0]

    ctx = dict(dist=dist, rank=rank, world=world, local=local, device=device, barrier=barrier, maxreduce=maxreduce, flush=flush)

    # ------------------------------------------------------------------ the headline workload
    main_run = Runner(ctx, args.workload, nph)
    p = main_run.p
    # the sampler starts BEFORE the warm-up: nvidia-smi takes about a second to initialise NVML over all GPUs of the box,
    # and that start-up (not the 200 ms sampling itself) was measured to cost the first timed steps ~60 ms
    sampler = ClockSampler(local)
    sampler.start()
    dev = main_run.device_steps(args.warmup, args.steps, sampler)
    e2e_total, e2e_out = main_run.e2e_steps(2, args.steps)
    total_ms, kern_ms, clocks, res = dev["total_ms"], dev["kernel_ms"], dev["clocks"], dev["res"]

    # ------------------------------------------------------------------ extra workloads (same line, `extra`)
    extra = {}
    for name in [x for x in args.extra.split(",") if x and x != args.workload]:
        xr = Runner(ctx, name, int(args.extra_photons))
        xd = xr.device_steps(1, args.extra_steps)
        xt, xo = xr.e2e_steps(1, args.extra_steps)
        if rank == 0:
            n = int(args.extra_photons)
            extra[name] = dict(photons_per_gpu=n, steps=args.extra_steps, warmup=1, value=n * world * args.extra_steps / xd["total_ms"], unit="photons/ms",
                               e2e=n * world * args.extra_steps / xt, kernel_ms=xd["kernel_ms"], ms_per_step=xd["total_ms"] / args.extra_steps,
                               absorbed=xd["res"]["absorbed"], detected=xd["res"]["detected"], kernel="photon_kernel<%s>" % xd["kname"],
                               parity_check=xr.parity_check(xd, xo))
            kc = kernel_counters(name)
            if kc:
                f_sm = ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6
                extra[name]["issue_frac"] = kc["warp_inst_per_photon"] * n / (xd["kernel_ms"] * 1e-3) / (N_SM * SCHEDULERS_PER_SM * f_sm)

    if rank == 0:
        nthread, kname, copies = dev["nthread"], dev["kname"], dev["copies"]
        value = nph * world * args.steps / total_ms
        out = e2e_out
        h2d = p.dimxyz * (1 if int((p.keep["vol"] & 0x7FFFFFFF).max()) < 128 else 2) + 16 * (p.c.medianum + 4 * (1 + p.c.extrasrclen) + p.c.detnum) + 16 * nthread
        d2h = 4 * p.fieldlen + 4 * p.reclen * int(out["saved"] if out else 0) + 16 + 4
        # roofline pieces
        w = WORK.get(args.workload)
        peak, peak_src = measured_peaks()
        roof = None
        if w:
            acc_bytes = 8
            alg_bytes = (w["seg"] * 1 + w["dep"] * acc_bytes) * nph
            hbm_ach = alg_bytes / (kern_ms * 1e-3) / 1e9
            ms_, ops_ = C.c_float(), C.c_uint64()
            abi.check(lib.mcxb_bench_red(local, 8, max(p.fieldlen, 1 << 16), 148 * 8, 2000, 0, 3, C.byref(ms_), C.byref(ops_)))
            red_peak = ops_.value / ms_.value / 1e6
            red_ach = w["dep"] * nph / (kern_ms * 1e-3) / 1e9
            hbm = dict(achieved=hbm_ach, peak=peak, unit="GB/s", frac=hbm_ach / peak, peak_source=peak_src,
                       algorithmic_bytes_per_photon=w["seg"] + w["dep"] * acc_bytes)
            l2 = dict(achieved=red_ach, peak=red_peak, unit="G reductions/s", frac=red_ach / red_peak,
                      peak_source="mcxb_bench_red: uniform random fp64 RED over a buffer of the volume's size, measured in this run")
            note = ("working set (media %.2f MB + fp64 accumulators %.2f MB) is L2-resident: DRAM traffic is a small fraction of the algorithmic bytes and HBM is not "
                    "the binding ceiling (hbm_frac); the kernel is bound by SM issue slots, with L2 reductions behind that (l2_red_frac)" % (p.dimxyz / 1e6, 8 * p.fieldlen / 1e6))
            kc = kernel_counters(args.workload)
            if kc:
                # the binding ceiling: warp instructions issued per second against 148 SMs x 4 schedulers x the SM clock
                f_sm = (clocks.get("sm_mhz") or 1965.0) * 1e6
                issue_peak = N_SM * SCHEDULERS_PER_SM * f_sm
                issue_ach = kc["warp_inst_per_photon"] * nph / (kern_ms * 1e-3)
                frac = issue_ach / issue_peak
                roof = dict(bound="issue", achieved=issue_ach / 1e12, peak=issue_peak / 1e12, unit="T warp-instructions/s", frac=frac,
                            traffic=kc["dram_bytes_per_photon"] * nph,
                            lane_occupancy=kc["lanes_per_warp_inst"] / 32.0, lanes_per_warp_inst=kc["lanes_per_warp_inst"],
                            thread_issue_frac=frac * kc["lanes_per_warp_inst"] / 32.0,
                            warp_inst_per_photon=kc["warp_inst_per_photon"],
                            peak_source="148 SMs x 4 schedulers x SM clock sampled during this run (%.0f MHz)" % (f_sm / 1e6),
                            counter_source=kc["source"], kernel="photon_kernel<%s>" % kname, kernel_ms=kern_ms,
                            hbm_frac=hbm["frac"], hbm_achieved_gbs=hbm_ach, hbm_peak_gbs=peak, l2_red_frac=l2["frac"], l2_red_achieved_gops=red_ach, l2_red_peak_gops=red_peak,
                            hbm=hbm, l2_red=l2, note=note)
            else:
                roof = dict(bound="hbm", traffic=None, kernel="photon_kernel<%s>" % kname, kernel_ms=kern_ms, l2_red=l2, note=note, **hbm)
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            info = bounded_cpu_run(args.workload, args.cpu_photons)
            cpu = dict(value=info["value"], unit="photons/ms", cores=info["cores"], kind=info["kind"],
                       sample="%d photons of %s, %d work-items, %.1f s" % (info["photons"], args.workload, info["work_items"], info["wall_ms"] / 1e3),
                       absorbed=info["absorbed"])
        line = dict(metric="photons/ms", value=value, unit="photons/ms", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=value / PUBLISHED_PHOTONS_PER_MS,
                    dtype="f32", data="synthetic",
                    config=dict(workload=args.workload, nphoton_per_gpu=nph, volume="%dx%dx%d" % p.dims, media="u8 labels", accumulators="fp64 RED, %d copies summed by finalize" % copies,
                                nthread=nthread, block=256, scheduling="dynamic photon counter", l2="flushed between steps (384 MB memset)",
                                parallelism="photon shards x%d + NCCL reduce/gather" % world if world > 1 else "single GPU",
                                baseline_ref="README.md:618: 12953.37 photon/ms on a Titan V-class GPU (OpenCL), 1e7 photons",
                                absorbed=res["absorbed"], detected=res["detected"], maxdetphoton=int(p.c.maxdetphoton)),
                    clocks=clocks,
                    e2e=dict(value=nph * world * args.steps / e2e_total, unit="photons/ms", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), ms_per_step=e2e_total / args.steps),
                    gpu_launches=dev["launches"], kernel_ms=kern_ms, kernel_photons_per_ms=nph / kern_ms,
                    kernel_ms_steps=dev["kernel_ms_steps"], parity_check=main_run.parity_check(dev, e2e_out))
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
