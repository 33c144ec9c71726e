#!/usr/bin/env python
"""bench.py -- particle-steps/s of the W8, 4th-order-composition symplectic PIC step.

    python bench.py --gpus N --steps K --warmup W            (ours: hand-written sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K --warmup W   (the reference's CPU code)

A "step" is one `Theta_map4` (include/strugepic_propagators.hpp:574-583 of the reference:
3 x map2 = 18 axis sub-flows + 6 Theta_E + 3 Theta_B) over a uniform thermal plasma.

Workload.  BASELINE.json quotes the metric on `examples/full: 512^3 cells x 64 ppc`, which
is 8.59e9 particles = 412 GB of FP64 phase space and does not fit one B200 (180 GB).  The
per-GPU brick measured here is the largest same-physics case that does: 256^3 cells x 64 ppc
= 1.07e9 particles (51.5 GB SoA, 64 GB with bin slack), W8 (P8R2), dt = 0.5, v_th = 0.01,
q = -1, m = 100, E = 0, B = (0,0,1), fully periodic.  With N GPUs the
global box is 256 x 256 x (256 N) cut into z slabs (weak scaling).

Keys: `value` = particle-steps/s with the state resident in HBM, CUDA-event timed on the
library's stream, max over ranks; `e2e` = the same step driven through the C ABI with HOST
buffers: fields and particles uploaded from pinned host memory (spic_set_field /
spic_set_particles), one spic_map(4), fields and particles read back (spic_get_*), all inside
the timed region; `roofline` = the dominant kernel (theta_axis, 18 of the 24 particle launches
of a step) in algorithmic bytes (72 B per particle per sub-flow) over its own CUDA-event
duration against MEASURED_PEAKS.json; `roofline_fp64` = the same launch in algorithmic FP64
flops (718 per particle per sub-flow) against the DFMA rate measured here by
spic_probe_fp64_tflops -- W8 is FP64-pipe bound, so that is the binding roof;
`cpu_baseline` / `--impl reference` = the reference's own sources (oracle/_ref, compiled
unmodified; else the C port) on every host core, one independent 16^3 x 64 ppc brick per
core (the communication-free upper bound of its MPI build).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle-steps/s (W8, 4th-order split)"
UNIT = "particle-steps/s"
BYTES_PER_SUBFLOW = 72.0        # R 6 doubles, W 3 doubles per particle per sub-flow
FLOP_AXIS, FLOP_PUSHVE = 718.0, 842.0   # SURVEY 8d, W8, factorised count
FLOP_AXIS_PWL, FLOP_PUSHVE_PWL = 80.0, 100.0
SUBFLOWS_PER_STEP = 24          # map4: 18 axis + 6 push_V_E
CPU_BRICK, CPU_PPC = 16, 64     # per-core sample of the reference arm


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", dest="n", type=int, default=256, help="cells per side of one GPU's brick")
    ap.add_argument("--ppc", type=int, default=64)
    ap.add_argument("--interp", default="p8r2", choices=["p8r2", "pwl"])
    ap.add_argument("--order", type=int, default=4, choices=[1, 2, 4])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--clock-period-ms", type=int, default=200,
                    help="nvidia-smi sampling period during the timed region (0: no sampler; diagnosis only)")
    ap.add_argument("--no-fuse", action="store_true", help="reference launch-per-sub-flow schedule (A/B against the fused axis block)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="library option for A/B runs (spic_set_option), e.g. --opt block_stream=1")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--cores", type=int, default=0, help="reference arm: worker processes (0 = all)")
    return ap.parse_args()


def workload_name(a, n_gpus):
    return ("uniform plasma %dx%dx%d cells, %d ppc, %s, Theta_map%d, dt=0.5, v_th=0.01 "
            "(stand-in for examples/full 512^3 x 64 ppc = 412 GB, which does not fit one B200)"
            % (a.n, a.n, a.n * n_gpus, a.ppc, "W8/P8R2" if a.interp == "p8r2" else "PWL", a.order))


# ------------------------------------------------------------------------------------------
# reference arm: the reference's CPU implementation on every host core
# ------------------------------------------------------------------------------------------
def _cpu_worker(conn, interp, order, seed):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import oracle as ora  # the reference's own code (oracle/_ref) or its C port: CPU baseline only
    from strugepic_b200 import synthetic
    n_cell = (CPU_BRICK,) * 3
    try:
        o = ora.RefOracle(n_cell, interp=interp, ng=2 if interp == 0 else 1, fast=True) \
            if ora.have_ref(interp, fast=True) else ora.PortOracle(n_cell, interp=interp, fast=True)
    except Exception as e:  # noqa: BLE001
        conn.send(("error", repr(e)))
        return
    E = np.zeros((3,) + n_cell)
    B = np.zeros((3,) + n_cell)
    B[2] = 1.0
    o.set_field(0, E)
    o.set_field(1, B)
    parts = synthetic.uniform_plasma(n_cell, CPU_PPC, 0.01, seed)
    o.set_particles(*parts, -1.0 / CPU_PPC, 100.0 / CPU_PPC)
    conn.send(("ready", o.kind, o.num_particles()))
    while True:
        cmd = conn.recv()
        if cmd == "stop":
            break
        t0 = time.perf_counter()
        o.map(order, 0.5)
        conn.send(time.perf_counter() - t0)


def run_reference(a, steps, warmup, cores=0):
    """Times `steps` Theta_map steps of the reference's CPU code on `cores` processes."""
    import multiprocessing as mp
    interp = 0 if a.interp == "p8r2" else 1
    ncores = cores or len(os.sched_getaffinity(0))
    ctx = mp.get_context("fork")
    procs = []
    for r in range(ncores):
        pc, cc = ctx.Pipe()
        p = ctx.Process(target=_cpu_worker, args=(cc, interp, a.order, 1000 + r), daemon=True)
        p.start()
        procs.append((p, pc))
    kind, npart = "port", 0
    for _, pc in procs:
        msg = pc.recv()
        if msg[0] == "error":
            raise RuntimeError("reference worker failed: " + msg[1])
        kind, npart = msg[1], npart + msg[2]
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        for _, pc in procs:
            pc.send("step")
        for _, pc in procs:
            pc.recv()
        dt = time.perf_counter() - t0  # max over the workers: the step ends when the last core ends
        if s >= warmup:
            times.append(dt)
    for p, pc in procs:
        pc.send("stop")
    for p, _ in procs:
        p.join(timeout=10)
    total = sum(times)
    return {"value": npart * steps / total, "ms_per_step": 1e3 * total / steps, "cores": ncores,
            "kind": "reference" if kind == "reference" else "port", "particles": npart,
            "sample": "%d independent %d^3 x %d ppc bricks (one per host core, %d particles in all), "
                      "%d Theta_map%d steps after %d warm-up, %s" %
                      (ncores, CPU_BRICK, CPU_PPC, npart, steps, a.order, warmup,
                       "reference sources compiled unmodified (oracle/_ref, -O3 -mavx2 -mfma)"
                       if kind == "reference" else "C port of the reference (oracle/port, -O3 -mavx2 -mfma)")}


def reference_main(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    r = run_reference(a, a.steps, a.warmup, a.cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a, a.gpus), "sample": r["sample"]},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device, period_ms=200):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        if period_ms <= 0:
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(device), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", str(period_ms)],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.f.read().splitlines():
            t = [x.strip() for x in ln.split(",")]
            if len(t) < 8:
                continue
            try:
                sm.append(float(t[0]))
                mx.append(float(t[1]))
                power.append(float(t[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), t[4:8]):
                if v.lower() == "active":
                    reasons.add(name)
        self.f.close()
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "power_w_max": max(power), "samples": len(sm)}
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per theta_axis launch from the committed ncu --set full capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(p):
        return json.load(open(p))
    return None


def ours_main(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        # reference CPU code timed beside us, BEFORE CUDA is initialised (fork-safe), rank 0, N=1 only
        try:
            cpu = run_reference(a, a.cpu_steps, 1, a.cores)
        except Exception as e:  # noqa: BLE001
            cpu = {"error": repr(e)}

    import numpy as np
    import torch
    import strugepic_b200 as spic  # raises ImportError when the CUDA library is missing: no fallback

    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def _noop():
        pass

    def _barrier():
        dist.barrier()

    barrier = _noop if dist is None else _barrier

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    interp = spic.P8R2 if a.interp == "p8r2" else spic.PWL
    n_cell = (a.n, a.n, a.n * world)
    sim = spic.Simulation(n_cell, interp=interp, device=local, nranks=world, rank=rank)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.tensor(list(spic.comm_unique_id()), dtype=torch.uint8, device="cuda")
        dist.broadcast(uid, 0)
        sim.comm_init(bytes(uid.cpu().tolist()))
    sim.set_uniform_field(spic.FIELD_E, [0.0, 0.0, 0.0])
    sim.set_uniform_field(spic.FIELD_B, [0.0, 0.0, 1.0])
    sim.add_particle_density_uniform(a.ppc, 100.0, -1.0, 0.01, seed=12345)
    sim.sync()
    npart_local = sim.num_particles()
    npart = int(allsum(float(npart_local)))
    fp64_peak = spic.probe_fp64_tflops(local, 0.5)
    fp64_peak3 = spic.probe_fp64_tflops(local, 0.25, three_operands=True)

    if a.no_fuse:
        sim.set_option("fuse", 0)
    for kv in a.opt:
        k, v = kv.split("=")
        sim.set_option(k, float(v))
    stream = torch.cuda.ExternalStream(sim.stream())
    for _ in range(a.warmup):
        sim.map(a.order, 0.5)
    sim.sync()

    # ---- timed region: K steps, state resident in HBM --------------------------------------
    sim.set_option("time_kernels", 1)
    sim.kernel_times(reset=True)
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local, a.clock_period_ms)
    torch.cuda.synchronize()
    barrier()
    e0.record(stream)
    for _ in range(a.steps):
        sim.map(a.order, 0.5)
    e1.record(stream)
    sim.sync()
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    ms = allmax(e0.elapsed_time(e1))
    launches = sim.launch_count() - l0
    kt = sim.kernel_times(reset=True)
    sim.set_option("time_kernels", 0)
    value = npart * a.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    # fused schedule (default on periodic single-rank boxes): k_axis_block = the six Theta of one map2 in one
    # launch; otherwise k_theta_axis_* = one Theta per launch.  Algorithmic work is counted on the REFERENCE
    # schedule either way (72 B and 718 flop per particle and reference sub-flow), so fusion shows
    # up as time saved, not as work invented; the fused lower bound (96 B per particle and block) is given too.
    peaks, peak_src = measured_peaks()
    fused = kt["axis_block"][1] > 0
    ax_ms, ax_n = kt["axis_block"] if fused else kt["theta_axis"]
    sub_per_launch = 6 if fused else 1
    pv_ms, pv_n = kt["push_V_E"]
    ax_avg = ax_ms / max(ax_n, 1)
    f_axis = FLOP_AXIS if a.interp == "p8r2" else FLOP_AXIS_PWL
    f_pv = FLOP_PUSHVE if a.interp == "p8r2" else FLOP_PUSHVE_PWL
    kname = "k_axis_block" if fused else ("k_theta_axis_v2" if a.ppc >= 40 else "k_theta_axis_v3")
    alg_bytes = BYTES_PER_SUBFLOW * sub_per_launch * npart_local
    achieved = alg_bytes / (ax_avg * 1e-3) / 1e9 if ax_n else 0.0
    traffic = ncu_traffic()
    tkey = "axis_block_dram_bytes_per_particle" if fused else "theta_axis_dram_bytes_per_particle"
    roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic[tkey] * npart_local if traffic and tkey in traffic else None,
                "traffic_source": (traffic["axis_block_capture" if fused else "capture"] +
                                   "; DRAM read+write bytes per particle x the particles of one launch here")
                if traffic and tkey in traffic else None,
                "peak_source": peak_src, "avg_launch_ms": ax_avg, "launches_timed": ax_n,
                "reference_subflows_per_launch": sub_per_launch,
                "algorithmic_bytes_per_launch": alg_bytes,
                "fused_lower_bound_bytes_per_launch": 96.0 * npart_local if fused else None,
                "share_of_step": ax_ms / ms if ms else None,
                "note": "W8 is FP64-pipe bound (10 flop/B > ridge 5.8): see roofline_fp64 for the binding roof"}
    ach_tf = f_axis * sub_per_launch * npart_local / (ax_avg * 1e-3) / 1e12 if ax_n else 0.0
    step_tf = (18 * f_axis + 6 * f_pv) * npart_local * a.steps / (ms * 1e-3) / 1e12 if a.order == 4 else None
    roofline_fp64 = {"kernel": kname, "bound": "fp64", "achieved": ach_tf, "peak": fp64_peak,
                     "unit": "TFLOP/s", "frac": ach_tf / fp64_peak if fp64_peak else None,
                     "peak_source": "measured here: spic_probe_fp64_tflops (dependent-free DFMA chains, all SMs)",
                     "three_register_operand_dfma_tflops": fp64_peak3,
                     "three_register_operand_note": "DFMA with three distinct register operands (gathers, deposition: "
                                                    "~2/3 of this kernel's FP64 instructions) issues every 3 cycles "
                                                    "instead of 2: the register file, not the pipe, bounds it",
                     "algorithmic_flop_per_particle": f_axis * sub_per_launch,
                     "flop_accounting": "reference schedule: 18 Theta x 718 + 6 Theta_E x 842 flop per particle-step "
                                        "(SURVEY 8d); the fused schedule executes fewer (shared weight evaluations, "
                                        "merged Theta_z and Theta_E halves)",
                     "whole_step_tflops": step_tf, "whole_step_frac": step_tf / fp64_peak if step_tf else None,
                     "push_V_E_avg_ms": pv_ms / max(pv_n, 1), "push_V_E_launches_per_step": pv_n / max(a.steps, 1),
                     "push_V_E_tflops": f_pv * npart_local / (pv_ms / max(pv_n, 1) * 1e-3) / 1e12 if pv_n else None}

    energy = sim.get_total_energy()
    # ---- e2e: the same step through the C ABI with HOST buffers ------------------------------
    e2e = None
    if not a.no_e2e:
        # pinned host memory for every local rank's brick must fit the host; otherwise the e2e leg
        # runs on a smaller brick per GPU (and says so)
        avail = 0
        try:
            for ln in open("/proc/meminfo"):
                if ln.startswith("MemAvailable"):
                    avail = int(ln.split()[1]) * 1024
        except OSError:
            pass
        e2e_n = a.n
        while avail and e2e_n > 32 and (48 * a.ppc + 48) * e2e_n ** 3 * world * 1.4 > avail:
            e2e_n //= 2
        e2e_n = int(-allmax(-float(e2e_n)))  # every rank must take the same decision
        if e2e_n != a.n:
            sim.close()
            sim = spic.Simulation((e2e_n, e2e_n, e2e_n * world), interp=interp, device=local, nranks=world, rank=rank)
            if world > 1:
                uid2 = torch.zeros(128, dtype=torch.uint8, device="cuda")
                if rank == 0:
                    uid2 = torch.tensor(list(spic.comm_unique_id()), dtype=torch.uint8, device="cuda")
                dist.broadcast(uid2, 0)
                sim.comm_init(bytes(uid2.cpu().tolist()))
            sim.set_uniform_field(spic.FIELD_E, [0.0, 0.0, 0.0])
            sim.set_uniform_field(spic.FIELD_B, [0.0, 0.0, 1.0])
            sim.add_particle_density_uniform(a.ppc, 100.0, -1.0, 0.01, seed=12345)
            sim.map(a.order, 0.5)
            sim.sync()
        n_loc = sim.num_particles()
        e2e = run_e2e(a, sim, spic, np, torch, n_loc, int(allsum(float(n_loc))), barrier, allmax, e2e_n)
        if e2e_n != a.n:
            e2e["workload"] = ("%dx%dx%d cells x %d ppc (per-GPU brick reduced from %d^3: pinned host buffers for "
                               "%d ranks would not fit %.0f GB of host RAM)" %
                               (e2e_n, e2e_n, e2e_n * world, a.ppc, a.n, world, avail / 1e9))

    sim.close()
    if rank == 0:
        line = {
            "metric": METRIC if a.interp == "p8r2" and a.order == 4 else METRIC.replace("W8", "PWL" if a.interp == "pwl" else "W8").replace("4th-order split", "Theta_map%d" % a.order if a.order != 4 else "4th-order split"),
            "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a, world), "particles": npart, "cells": a.n ** 3 * world,
                       "parallelism": "z-slab x%d" % world,
                       "l2": "inputs (%.1f GB of particle state per GPU) exceed the 126 MB L2; no flush needed"
                             % (48e-9 * npart_local)},
            "schedule": ("fused: per map2 one axis-block launch (x y z z y x) + Theta_B, adjacent Theta_E halves merged"
                         if fused else "reference: one launch per sub-flow"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "roofline_fp64": roofline_fp64,
            "cell_updates_per_s": a.n ** 3 * world * a.steps / (ms * 1e-3),
            "kernel_ms_per_step": {k: v[0] / a.steps for k, v in kt.items()},
            "energy_after": energy,
        }
        if cpu is not None and "error" not in cpu:
            line["cpu_baseline"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "kind": cpu["kind"],
                                    "sample": cpu["sample"]}
        elif cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(a, sim, spic, np, torch, npart_local, npart, barrier, allmax, cells):
    """Upload (pinned host -> device, re-bin), one Theta_map, read back: all timed."""
    # with N > 1 particles migrate between slabs, so the per-rank count changes from step to step:
    # pinned buffers carry 2 % slack and every transfer uses the live count
    cap = int(npart_local * 1.02) + 4096 if barrier.__name__ != "_noop" else npart_local
    pinned = [torch.empty(cap, dtype=torch.float64, pin_memory=True).numpy() for _ in range(6)]
    host_f = [torch.empty((3, cells, cells, cells), dtype=torch.float64, pin_memory=True).numpy() for _ in range(2)]
    n_live = sim.num_particles()
    sim.get_particles(0, out=[t[:n_live] for t in pinned])
    sim.get_field(spic.FIELD_E, out=host_f[0])
    sim.get_field(spic.FIELD_B, out=host_f[1])
    sim.sync()
    fbytes = sum(t.nbytes for t in host_f)
    h2d = d2h = 0
    times = []
    for s in range(1 + a.e2e_steps):
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        sim.set_field(spic.FIELD_E, host_f[0])
        sim.set_field(spic.FIELD_B, host_f[1])
        t1 = time.perf_counter()
        sim.set_particles(0, *[t[:n_live] for t in pinned])
        sim.sync()
        t2 = time.perf_counter()
        up = 48 * n_live + fbytes
        sim.map(a.order, 0.5)
        sim.sync()
        t3 = time.perf_counter()
        sim.get_field(spic.FIELD_E, out=host_f[0])
        sim.get_field(spic.FIELD_B, out=host_f[1])
        sim.sync()
        t4 = time.perf_counter()
        n_live = sim.num_particles()
        if n_live > cap:
            raise RuntimeError("e2e: pinned particle buffer too small after migration")
        sim.get_particles(0, out=[t[:n_live] for t in pinned])
        sim.sync()
        torch.cuda.synchronize()
        phases = {"set_fields": t1 - t0, "set_particles": t2 - t1, "map": t3 - t2, "get_fields": t4 - t3,
                  "get_particles": time.perf_counter() - t4}
        dt = time.perf_counter() - t0
        barrier()
        if s >= 1:
            times.append(allmax(dt))
            h2d, d2h = up, 48 * n_live + fbytes
    t = sum(times) / len(times)
    return {"value": npart / t, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "bytes_are": "per rank",
            "ms_per_step": 1e3 * t, "steps": a.e2e_steps,
            "phase_ms_last_step_rank0": {k: round(1e3 * v, 1) for k, v in phases.items()},
            "path": "spic_set_field x2 + spic_set_particles (pinned host -> HBM, re-binned) + spic_map + "
                    "spic_get_field x2 + spic_get_particles (HBM -> pinned host), wall clock, max over ranks"}


if __name__ == "__main__":
    args = parse()
    sys.exit(reference_main(args) if args.impl == "reference" else ours_main(args))
