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
library's stream (the closing event after the spic_sync that applies the deferred half kick),
max over ranks; `e2e` = the same step driven through the C ABI with HOST buffers: fields and
particles uploaded from pinned host memory (spic_set_field / spic_set_particles), one
spic_map(4), fields and particles read back (spic_get_*), all inside the timed region;
`checks` = particle count, discrete Gauss residual drift and energy before / after the timed
region (no oracle needed; collective over the slabs at N > 1); `roofline` = the dominant
kernel (k_axis_block: one axis block = the six Theta of a map2) in algorithmic bytes of the
REFERENCE schedule (6 x 72 B per particle) over its own CUDA-event duration against
MEASURED_PEAKS.json, with the fused byte bound beside it; `roofline_fp64` = the same block in
algorithmic FP64 flops (6 x 718 per particle) against the better of two DFMA probes
measured here -- W8 is FP64-pipe bound, so that is the binding roof; `secondary` (N = 1) =
512^3 x 8 ppc, the PWL variant and the vacuum field_only step of BASELINE configs[2];
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
                    help="library option for A/B runs (spic_set_option), e.g. --opt fuse=0")
    ap.add_argument("--walls", action="store_true",
                    help="x walls (MABC + reflecting particles, as examples/full/bernstein_main.cpp): the plasma stays "
                         "W + 2 cells away from them; A/B of the half-block schedule with --no-fuse")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary records (low ppc, PWL, field_only)")
    ap.add_argument("--secondary-all-ranks", action="store_true", help="run the secondary records at N > 1 too")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--cores", type=int, default=0, help="reference arm: worker processes (0 = all)")
    return ap.parse_args()


def workload_name(a, n_gpus):
    return ("uniform plasma %dx%dx%d cells%s, %d ppc, %s, Theta_map%d, dt=0.5, v_th=0.01 "
            "(stand-in for examples/full 512^3 x 64 ppc = 412 GB, which does not fit one B200)"
            % (a.n, a.n, a.n * n_gpus, " with x walls (MABC, reflection)" if a.walls else "", a.ppc,
               "W8/P8R2" if a.interp == "p8r2" else "PWL", a.order))


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


class Env:
    """Process-wide plumbing of our arm: torch.distributed rendezvous, reductions over ranks."""

    def __init__(self):
        import numpy as np
        import torch
        import strugepic_b200 as spic  # raises ImportError when the CUDA library is missing: no fallback
        self.np, self.torch, self.spic = np, torch, spic
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()

    def _red(self, x, op):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def allmax(self, x):
        return self._red(x, self.dist.ReduceOp.MAX) if self.dist else x

    def allsum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM) if self.dist else x

    def make_sim(self, n, interp, ppc, periodic=(1, 1, 1), opts=(), slab_profile=False):
        """One brick of n^3 cells per GPU (z slabs), E = 0, B = (0,0,1), uniform thermal plasma (ppc = 0: vacuum)."""
        spic, torch = self.spic, self.torch
        sim = spic.Simulation((n, n, n * self.world), periodic=periodic, interp=interp, device=self.local,
                              nranks=self.world, rank=self.rank)
        if self.world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if self.rank == 0:
                uid = torch.tensor(list(spic.comm_unique_id()), dtype=torch.uint8, device="cuda")
            self.dist.broadcast(uid, 0)
            sim.comm_init(bytes(uid.cpu().tolist()))
        sim.set_uniform_field(spic.FIELD_E, [0.0, 0.0, 0.0])
        sim.set_uniform_field(spic.FIELD_B, [0.0, 0.0, 1.0] if ppc else [0.0, 0.0, 0.0])
        if ppc and slab_profile:  # nothing within W + 2 cells of an x wall (the reflect cells)
            margin = sim.W + 2
            sim.add_particle_density(lambda nc, i, j, k: ((i >= margin) & (i < nc[0] - margin)) * 1.0 + 0.0 * (j + k),
                                     ppc, 100.0, -1.0, 0.01, seed=12345)
        elif ppc:
            sim.add_particle_density_uniform(ppc, 100.0, -1.0, 0.01, seed=12345)
        for kv in opts:
            k, v = kv.split("=")
            sim.set_option(k, float(v))
        sim.sync()
        return sim

    def invariants(self, sim, x_walls=False):
        """What the reader can check without an oracle: particle count, discrete Gauss residual, energy."""
        n = int(self.allsum(float(sim.num_particles()))) if sim.num_species() else 0
        g = sim.gauss_residual()  # collective over the slabs
        if x_walls:  # MABC blends E on the two face planes (hpp:447-476): div E moves there by construction
            g = g[:, :, 2:-2]
        return {"n": n, "g": g, "h": sum(sim.get_total_energy())}

    def checks(self, before, after):
        np = self.np
        gmax = self.allmax(float(np.max(np.abs(before["g"]))))
        drift = self.allmax(float(np.max(np.abs(after["g"] - before["g"]))))
        return {"particles_before": before["n"], "particles_after": after["n"],
                "particles_conserved": before["n"] == after["n"],
                "gauss_residual_max": gmax, "gauss_drift_max": drift,
                "gauss_ok": bool(drift < 1e-12 * max(1.0, gmax)),
                "energy_before": before["h"], "energy_after": after["h"],
                "energy_rel_change": (after["h"] - before["h"]) / before["h"] if before["h"] else 0.0,
                "what": "taken outside the timed region, before and after it: global particle count, max over cells "
                        "and ranks of |G1 - G0| with G = div E - rho (spic_gauss_residual: charge conservation of the "
                        "deposition, north_star check 2: < 1e-12 max(1,|G|)), H = field + kinetic energy (spic_energy)"}


def timed_steps(env, sim, step, steps, warmup, clock_period_ms=0, x_walls=False):
    """W untimed + K timed calls of step() on the library's stream: CUDA events, barrier + synchronize on both sides,
    max over ranks.  The last event is recorded after spic_sync, which applies the deferred half kick of the last
    step: every launch of the K steps is inside the clock."""
    torch = env.torch
    stream = torch.cuda.ExternalStream(sim.stream())
    for _ in range(warmup):
        step()
    sim.sync()
    before = env.invariants(sim, x_walls)
    sim.set_option("time_kernels", 1)
    sim.kernel_times(reset=True)
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(env.local, clock_period_ms)
    torch.cuda.synchronize()
    env.barrier()
    e0.record(stream)
    for _ in range(steps):
        step()
    sim.sync()
    e1.record(stream)
    torch.cuda.synchronize()
    env.barrier()
    clocks = sampler.stop()
    ms = env.allmax(e0.elapsed_time(e1))
    launches = sim.launch_count() - l0
    kt = sim.kernel_times(reset=True)
    sim.set_option("time_kernels", 0)
    after = env.invariants(sim, x_walls)
    return {"ms": ms, "launches": launches, "kt": kt, "clocks": clocks, "checks": env.checks(before, after)}


def fp64_roof(env):
    """FP64 denominator: the better of the two dependent-free DFMA probes, named."""
    spic = env.spic
    two = spic.probe_fp64_tflops(env.local, 0.5)
    imm = spic.probe_fp64_tflops(env.local, 0.5, immediate=True)
    three = spic.probe_fp64_tflops(env.local, 0.25, three_operands=True)
    return {"peak": max(two, imm), "two_constant_operand_dfma_tflops": two, "immediate_horner_dfma_tflops": imm,
            "three_register_operand_dfma_tflops": three, "nominal_tflops": 37.2,
            "peak_source": "measured here, max of spic_probe_fp64_immediate_tflops (Horner steps with immediate "
                           "coefficients) and spic_probe_fp64_tflops (two loop-constant operands): %s wins"
                           % ("immediate" if imm >= two else "two-constant")}


def particle_rooflines(a, env, r, npart_local, ppc, interp_name, order, steps, roof):
    """roofline / roofline_fp64 of the dominant kernel from the event pairs recorded inside the timed region."""
    kt, ms = r["kt"], r["ms"]
    peaks, peak_src = measured_peaks()
    fused = kt["axis_block"][1] > 0
    ax_ms, ax_n = kt["axis_block"] if fused else kt["theta_axis"]
    sub_per_launch = 6 if fused else 1
    # one "launch" of the fused schedule = one axis block = the six Theta of a map2; with z slabs a block is two
    # kernel launches (slab-face planes, then the interior planes): they are counted as one
    blocks = {4: 3, 2: 1}.get(order, 1) * steps
    n_units = blocks if fused else max(ax_n, 1)
    pv_ms, pv_n = kt["push_V_E"]
    ax_avg = ax_ms / max(n_units, 1)
    f_axis = FLOP_AXIS if interp_name == "p8r2" else FLOP_AXIS_PWL
    f_pv = FLOP_PUSHVE if interp_name == "p8r2" else FLOP_PUSHVE_PWL
    kname = "k_axis_block" if fused else ("k_theta_axis_v2" if ppc >= 40 else "k_theta_axis_v3")
    alg_bytes = BYTES_PER_SUBFLOW * sub_per_launch * npart_local
    achieved = alg_bytes / (ax_avg * 1e-3) / 1e9 if ax_n else 0.0
    traffic = ncu_traffic()
    tkey = "axis_block_dram_bytes_per_particle" if fused else "theta_axis_dram_bytes_per_particle"
    fused_bytes = 96.0 * npart_local if fused else None
    roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
                "traffic": traffic[tkey] * npart_local if traffic and tkey in traffic else None,
                "traffic_source": (traffic["axis_block_capture" if fused else "capture"] +
                                   "; DRAM read+write bytes per particle x the particles of one launch here")
                if traffic and tkey in traffic else None,
                "peak_source": peak_src, "avg_launch_ms": ax_avg, "launches_timed": ax_n,
                "kernel_launches_per_axis_block": ax_n / max(blocks, 1) if fused else None,
                "reference_subflows_per_launch": sub_per_launch,
                "algorithmic_bytes_per_launch": alg_bytes,
                "fused_lower_bound_bytes_per_launch": fused_bytes,
                "fused_bytes_achieved_gbs": fused_bytes / (ax_avg * 1e-3) / 1e9 if fused and ax_n else None,
                "fused_bytes_frac": fused_bytes / (ax_avg * 1e-3) / 1e9 / peaks["hbm_gbs"] if fused and ax_n else None,
                "share_of_step": ax_ms / ms if ms else None,
                # the whole step in contract bytes: 24 sub-flows x 72 B on the reference schedule, and the
                # lower bound once the sub-flows are fused by the exact commutations (3 x 96 + 4 x 72 B)
                "whole_step_reference_bytes_gbs": 1728.0 * npart_local * steps / (ms * 1e-3) / 1e9 if order == 4 else None,
                "whole_step_fused_bound_bytes_gbs": 576.0 * npart_local * steps / (ms * 1e-3) / 1e9 if order == 4 else None,
                "note": ("W8 is FP64-pipe bound (10 flop/B > ridge 5.8): see roofline_fp64 for the binding roof"
                         if interp_name == "p8r2" else
                         "PWL is HBM bound; `achieved` counts reference-schedule bytes (6 x 72 B per particle and "
                         "launch) on a kernel that moves them once: the binding figure is fused_bytes_frac")}
    ach_tf = f_axis * sub_per_launch * npart_local / (ax_avg * 1e-3) / 1e12 if ax_n else 0.0
    step_tf = (18 * f_axis + 6 * f_pv) * npart_local * steps / (ms * 1e-3) / 1e12 if order == 4 else None
    roofline_fp64 = {"kernel": kname, "bound": "fp64", "achieved": ach_tf, "peak": roof["peak"],
                     "unit": "TFLOP/s", "frac": ach_tf / roof["peak"] if roof["peak"] else None,
                     "frac_of_nominal": ach_tf / roof["nominal_tflops"],
                     "peak_source": roof["peak_source"],
                     "probes": {k: roof[k] for k in ("two_constant_operand_dfma_tflops", "immediate_horner_dfma_tflops",
                                                     "three_register_operand_dfma_tflops", "nominal_tflops")},
                     "three_register_operand_note": "DFMA with three distinct register operands (gathers, deposition: "
                                                    "~2/3 of this kernel's FP64 instructions) issues every 3 cycles "
                                                    "instead of 2: the register file, not the pipe, bounds it",
                     "algorithmic_flop_per_particle": f_axis * sub_per_launch,
                     "flop_accounting": "reference schedule: 18 Theta x 718 + 6 Theta_E x 842 flop per particle-step "
                                        "(SURVEY 8d); the fused schedule executes fewer (shared weight evaluations, "
                                        "merged Theta_z and Theta_E halves)",
                     "whole_step_tflops": step_tf, "whole_step_frac": step_tf / roof["peak"] if step_tf else None,
                     "push_V_E_avg_ms": pv_ms / max(pv_n, 1), "push_V_E_launches_per_step": pv_n / max(steps, 1),
                     "push_V_E_tflops": f_pv * npart_local / (pv_ms / max(pv_n, 1) * 1e-3) / 1e12 if pv_n else None}
    return fused, roofline, roofline_fp64


def secondary_particles(a, env, roof, name, n, ppc, interp_name):
    """A secondary particle shape (SURVEY 8d C4 / north_star): same step, same timing rules, short record."""
    spic = env.spic
    interp = spic.P8R2 if interp_name == "p8r2" else spic.PWL
    sim = env.make_sim(n, interp, ppc)
    npart_local = sim.num_particles()
    npart = int(env.allsum(float(npart_local)))
    steps = 2
    r = timed_steps(env, sim, lambda: sim.map(4, 0.5), steps, 3)
    _, rl, rl64 = particle_rooflines(a, env, r, npart_local, ppc, interp_name, 4, steps, roof)
    sim.close()
    keep = ("kernel", "bound", "achieved", "peak", "unit", "frac", "avg_launch_ms", "share_of_step")
    out = {"name": name, "metric": "particle-steps/s (%s, 4th-order split)" % ("W8" if interp_name == "p8r2" else "PWL"),
           "workload": "uniform plasma %d^3 cells x %d ppc per GPU, %s, Theta_map4, dt=0.5, v_th=0.01"
                       % (n, ppc, "W8/P8R2" if interp_name == "p8r2" else "PWL"),
           "value": npart * steps / (r["ms"] * 1e-3), "unit": UNIT, "particles": npart, "steps": steps, "warmup": 3,
           "ms_per_step": r["ms"] / steps, "gpu_launches": r["launches"],
           "kernel_ms_per_step": {k: v[0] / steps for k, v in r["kt"].items()},
           "roofline_fp64": {k: rl64[k] for k in keep if k in rl64} | {"frac_of_nominal": rl64["frac_of_nominal"],
                                                                        "whole_step_frac": rl64["whole_step_frac"]},
           "roofline": {k: rl[k] for k in keep if k in rl} | {"fused_bytes_achieved_gbs": rl["fused_bytes_achieved_gbs"],
                                                              "fused_bytes_frac": rl["fused_bytes_frac"]},
           "checks": r["checks"]}
    return out


def secondary_field_only(a, env, n=256):
    """BASELINE configs[2]: vacuum Maxwell n^3 with the soft plane-wave source and MABC in x, field sub-flows only
    (examples/field_only/main.cpp:142-145).  Algorithmic traffic: 3 curl sweeps x 72 B per cell and step."""
    spic = env.spic
    sim = env.make_sim(n, spic.P8R2, 0, periodic=(0, 1, 1))
    state = {"k": 0}

    def step():
        sim.field_only_step(4, 1, 0.1, 0.02, 0.5, state["k"])  # sp = 4, comp Y, Es = 0.1, omega = 0.02: source_absorb.input
        state["k"] += 1

    steps = 200
    r = timed_steps(env, sim, step, steps, 20)
    sim.close()
    peaks, peak_src = measured_peaks()
    cells = n ** 3 * env.world
    value = cells * steps / (r["ms"] * 1e-3)
    gbs = 216.0 * n ** 3 * steps / (r["ms"] * 1e-3) / 1e9
    return {"name": "field_only", "metric": "cell-updates/s (vacuum Maxwell, source + MABC-x)",
            "workload": "%d^3 cells per GPU, x walls (MABC) + soft plane-wave source, y z periodic, dt=0.5" % n,
            "value": value, "unit": "cell-updates/s", "steps": steps, "warmup": 20, "ms_per_step": r["ms"] / steps,
            "gpu_launches": r["launches"], "kernel_ms_per_step": {k: v[0] / steps for k, v in r["kt"].items()},
            "roofline": {"kernel": "field step (Theta_E/2, source, Theta_B, Theta_E/2: 3 curl sweeps)", "bound": "hbm",
                         "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                         "peak_source": peak_src, "algorithmic_bytes_per_cell_step": 216.0,
                         "roof_cell_updates_per_s": peaks["hbm_gbs"] * 1e9 / 216.0},
            "energy_after": r["checks"]["energy_after"]}


def ours_main(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu:
        # reference CPU code timed beside us, BEFORE CUDA is initialised (fork-safe), rank 0, N=1 only
        try:
            cpu = run_reference(a, a.cpu_steps, 1, a.cores)
        except Exception as e:  # noqa: BLE001
            cpu = {"error": repr(e)}

    env = Env()
    spic, np, torch, dist, local = env.spic, env.np, env.torch, env.dist, env.local
    interp = spic.P8R2 if a.interp == "p8r2" else spic.PWL
    opts = list(a.opt) + (["fuse=0"] if a.no_fuse else [])
    sim = env.make_sim(a.n, interp, a.ppc, periodic=(0, 1, 1) if a.walls else (1, 1, 1), opts=opts, slab_profile=a.walls)
    npart_local = sim.num_particles()
    npart = int(env.allsum(float(npart_local)))
    roof = fp64_roof(env)

    # ---- timed region: K steps, state resident in HBM --------------------------------------
    r = timed_steps(env, sim, lambda: sim.map(a.order, 0.5), a.steps, a.warmup, a.clock_period_ms, x_walls=a.walls)
    ms, launches, kt, clocks = r["ms"], r["launches"], r["kt"], r["clocks"]
    value = npart * a.steps / (ms * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    # fused schedule (default on periodic boxes): k_axis_block = the six Theta of one map2 in one launch; otherwise
    # k_theta_axis_* = one Theta per launch.  Algorithmic work is counted on the REFERENCE schedule either way (72 B
    # and 718 flop per particle and reference sub-flow), so fusion shows up as time saved, not as work
    # invented; the fused lower bound (96 B per particle and block) is given too.
    fused, roofline, roofline_fp64 = particle_rooflines(a, env, r, npart_local, a.ppc, a.interp, a.order, a.steps, roof)

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
        e2e_n = int(-env.allmax(-float(e2e_n)))  # every rank must take the same decision
        if e2e_n != a.n:
            sim.close()
            sim = env.make_sim(e2e_n, interp, a.ppc, opts=opts)
            sim.map(a.order, 0.5)
            sim.sync()
        n_loc = sim.num_particles()
        e2e = run_e2e(a, sim, spic, np, torch, n_loc, int(env.allsum(float(n_loc))), env.barrier if dist else None,
                      env.allmax, e2e_n)
        if e2e_n != a.n:
            e2e["workload"] = ("%dx%dx%d cells x %d ppc (per-GPU brick reduced from %d^3: pinned host buffers for "
                               "%d ranks would not fit %.0f GB of host RAM)" %
                               (e2e_n, e2e_n, e2e_n * world, a.ppc, a.n, world, avail / 1e9))
    sim.close()
    del sim

    # ---- secondary records (after the main region): the other shapes SURVEY 8d / BASELINE configs name ------------
    secondary = []
    if not a.no_secondary and (world == 1 or a.secondary_all_ranks) and a.interp == "p8r2" and a.order == 4:
        for fn in (lambda: secondary_particles(a, env, roof, "low_ppc", 2 * a.n, max(a.ppc // 8, 1), "p8r2"),
                   lambda: secondary_particles(a, env, roof, "pwl", a.n, a.ppc, "pwl"),
                   lambda: secondary_field_only(a, env, a.n)):
            try:
                secondary.append(fn())
            except Exception as e:  # noqa: BLE001  (a secondary must not take the headline down with it)
                secondary.append({"error": repr(e)[:300]})

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
            "schedule": ("fused: per map2 one axis block (x y z z y x) + Theta_B, adjacent Theta_E halves merged"
                         + ("; z slabs: slab-face planes first, their halo sums + migration on the comm stream under "
                            "the interior planes" if world > 1 else "")
                         if fused else "reference: one launch per sub-flow"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
            "roofline": roofline, "roofline_fp64": roofline_fp64,
            "checks": r["checks"],
            "cell_updates_per_s": a.n ** 3 * world * a.steps / (ms * 1e-3),
            "kernel_ms_per_step": {k: v[0] / a.steps for k, v in kt.items()},
            "energy_after": r["checks"]["energy_after"],
            "secondary": secondary,
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
    multi = barrier is not None
    if barrier is None:
        def barrier():
            pass
    cap = int(npart_local * 1.02) + 4096 if multi else npart_local
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
