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
  roofline       the contract's HBM roofline of the photon kernel (algorithmic bytes / kernel time vs the measured
                 copy bandwidth) plus the two ceilings that actually bound this L2-resident kernel: L2 reduction
                 throughput (measured live with a RED microbenchmark) and SM issue-slot use (from the committed
                 ncu capture, profiles/)
  cpu_baseline   the reference kernel source built for the host (oracle/_ref), all host cores, bounded sample

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


def bounded_cpu_run(workload, photons, hostthreads=0):
    """the reference kernel source on the host cores (oracle/_ref), or the C restatement when it is absent"""
    from mcxcl_b200 import benchmarks, hostcfg
    from oracle import loader
    kind = "reference" if loader.have_ref() else "port"
    chk = loader.ref() if kind == "reference" else loader.port()
    p = hostcfg.prepare(benchmarks.get(workload, photons))
    cores = os.cpu_count() or 1
    work = 64 * cores * 4
    t0 = time.perf_counter()
    o = chk.run(p, work, hostthreads=hostthreads)
    wall = (time.perf_counter() - t0) * 1e3
    return dict(kind=kind, cores=cores if hostthreads <= 0 else hostthreads, photons=int(photons), kernel_ms=o["runtime_ms"], wall_ms=wall,
                value=photons / o["runtime_ms"], absorbed=o["absorbed"])


def run_reference_arm(args, rank):
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
                config=dict(workload=args.workload, nphoton_per_step=photons, note="bounded sample of the workload; CPU implementation = the reference's kernel source (src/mcx_core.cl) compiled for the host through oracle/clshim.h, OpenMP over all cores"),
                cpu_baseline=dict(value=v, unit="photons/ms", cores=info["cores"], kind=info["kind"], sample="%d photons of %s per step" % (photons, args.workload)),
                e2e=dict(value=v, unit="photons/ms", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    from mcxcl_b200 import abi, benchmarks, engine, hostcfg, multigpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the photon-transport engine has no CPU path")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = abi.load()
    nph = int(args.photons)
    cfg = benchmarks.get(args.workload, nph)
    p = hostcfg.prepare(cfg)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident steps
    sim = engine.Simulation(p, local)
    sim.reseed(p.c.seed, rank * sim.nthread)
    # the L2-flush buffer is allocated AFTER the simulation's buffers: with the 384 MB block allocated first the driver
    # places the (2 MB) working set differently and the same kernel was measured 3.5 % slower (360.4 ms vs 348.0 ms,
    # profiles/README.md "placement"); a front-end that calls the engine has no such block in the way
    flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=device)      # > 126 MB of L2
    sampler = ClockSampler(local)
    kernel_ms, step_ms = [], []
    launches = 0

    def step(timed):
        nonlocal launches
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        sim.reset()
        sim.launch()
        if world > 1:
            multigpu.combine(sim, dist, rank, world, device)
        else:
            sim.finalize()
        e1.record()
        e1.synchronize()
        if timed:
            step_ms.append(e0.elapsed_time(e1))
            kernel_ms.append(sim.kernel_ms())
            launches += 2

    # the sampler starts BEFORE the warm-up: nvidia-smi takes about a second to initialise NVML over all GPUs of the box,
    # and that start-up (not the 200 ms sampling itself) was measured to cost the first timed steps ~60 ms
    sampler.start()
    for _ in range(args.warmup):
        step(False)
    barrier()
    sampler.rows.clear()                       # keep only what is sampled during the timed region
    for _ in range(args.steps):
        step(True)
    barrier()
    clocks = sampler.summary()
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=device)
    kern = torch.tensor([float(np.mean(kernel_ms))], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(kern, op=dist.ReduceOp.MAX)
    total_ms, kern_ms = float(total_ms.item()), float(kern.item())
    res = sim.fetch() if rank == 0 else None
    nthread, kname, copies = sim.nthread, sim.kernel_name, sim.acc_copies
    sim.close()

    # ------------------------------------------------------------------ end to end through the public host-buffer API
    e2e_ms = []
    p_all = p.clone_for(nphoton=nph * world)           # the whole job's budget; every rank takes its share of it
    for i in range(2 + args.steps):
        barrier()
        t0 = time.perf_counter()
        if world > 1:
            out = multigpu.run_distributed(p_all)
        else:
            out = engine.run_prepared(p, local)
        barrier()
        if i >= 2:
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
    e2e = torch.tensor([sum(e2e_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    e2e_total = float(e2e.item())

    if rank == 0:
        value = nph * world * args.steps / total_ms
        h2d = p.dimxyz * (1 if int((p.keep["vol"] & 0x7FFFFFFF).max()) < 128 else 2) + 16 * (p.c.medianum + 4 * (1 + p.c.extrasrclen) + p.c.detnum) + 16 * nthread
        d2h = 4 * p.fieldlen + 4 * p.reclen * int(out["saved"] if out else 0) + 16 + 4
        # roofline pieces
        w = WORK.get(args.workload)
        peak, peak_src = measured_peaks()
        roof = None
        if w:
            acc_bytes = 8
            alg_bytes = (w["seg"] * 1 + w["dep"] * acc_bytes) * nph
            achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
            ms_, ops_ = C.c_float(), C.c_uint64()
            abi.check(lib.mcxb_bench_red(local, 8, max(p.fieldlen, 1 << 16), 148 * 8, 2000, 0, 3, C.byref(ms_), C.byref(ops_)))
            red_peak = ops_.value / ms_.value / 1e6
            red_ach = w["dep"] * nph / (kern_ms * 1e-3) / 1e9
            kc = kernel_counters(args.workload)
            issue = None
            traffic = None
            if kc:
                f_sm = (clocks.get("sm_mhz") or 1965.0) * 1e6
                issue_peak = N_SM * SCHEDULERS_PER_SM * f_sm                       # warp instructions per second
                issue_ach = kc["warp_inst_per_photon"] * nph / (kern_ms * 1e-3)
                issue = dict(achieved=issue_ach / 1e12, peak=issue_peak / 1e12, unit="T warp-instructions/s", frac=issue_ach / issue_peak,
                             warp_inst_per_photon=kc["warp_inst_per_photon"], lanes_per_warp_inst=kc["lanes_per_warp_inst"],
                             peak_source="148 SMs x 4 schedulers x SM clock sampled during this run (%.0f MHz)" % (f_sm / 1e6),
                             counter_source=kc["source"])
                traffic = kc["dram_bytes_per_photon"] * nph
            roof = dict(bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, traffic=traffic, peak_source=peak_src,
                        kernel="photon_kernel<%s>" % kname, kernel_ms=kern_ms,
                        algorithmic_bytes_per_photon=w["seg"] + w["dep"] * acc_bytes,
                        note="working set (media %.2f MB + fp64 accumulators %.2f MB) is L2-resident: DRAM traffic is a small fraction of the algorithmic bytes and HBM is not the binding ceiling; the kernel is bound by SM issue slots (issue) and, behind that, by L2 reductions (l2_red)" % (p.dimxyz / 1e6, 8 * p.fieldlen / 1e6),
                        issue=issue,
                        l2_red=dict(achieved=red_ach, peak=red_peak, unit="G reductions/s", frac=red_ach / red_peak, peak_source="mcxb_bench_red: uniform random fp64 RED over a buffer of the volume's size, measured in this run"))
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            info = bounded_cpu_run(args.workload, args.cpu_photons)
            cpu = dict(value=info["value"], unit="photons/ms", cores=info["cores"], kind=info["kind"],
                       sample="%d photons of %s, %d work-items, %.1f s" % (info["photons"], args.workload, 64 * info["cores"] * 4, info["wall_ms"] / 1e3),
                       absorbed=info["absorbed"])
        line = dict(metric="photons/ms", value=value, unit="photons/ms", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=total_ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=value / PUBLISHED_PHOTONS_PER_MS,
                    dtype="f32", data="synthetic",
                    config=dict(workload=args.workload, nphoton_per_gpu=nph, volume="%dx%dx%d" % p.dims, media="u8 labels", accumulators="fp64 RED, %d copies summed by finalize" % copies,
                                nthread=nthread, block=256, scheduling="dynamic photon counter", l2="flushed between steps (384 MB memset)",
                                parallelism="photon shards x%d + NCCL reduce/gather" % world if world > 1 else "single GPU",
                                baseline_ref="README.md:618: 12953.37 photon/ms on a Titan V-class GPU (OpenCL), 1e7 photons",
                                absorbed=res["absorbed"], detected=res["detected"]),
                    clocks=clocks,
                    e2e=dict(value=nph * world * args.steps / e2e_total, unit="photons/ms", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h), ms_per_step=e2e_total / args.steps),
                    gpu_launches=launches, kernel_ms=kern_ms, kernel_photons_per_ms=nph / kern_ms,
                    kernel_ms_steps=[round(x, 2) for x in kernel_ms])
        if roof:
            line["roofline"] = roof
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
