#!/usr/bin/env python
"""bench.py - fp64 atom-steps/s of the Lennard-Jones integrator on B200 (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W  # the reference's CPU integrator (oracle/_ref)

A "step" is one LennardJonesSimulation::integrate() of the whole system.  Workload (configs[3] of
BASELINE.json, the one the metric is quoted on): N = 1,048,576 LJ atoms, rho* = 0.8, kT = 1, rc = 2.5,
shell = 0.4, tau = 0.5, dt = 0.001, cubic periodic box, the reference's own start lattice + Gaussian
velocities, equilibrated for --equil steps before the warm-up (SURVEY.md 8d protocol).  Synthetic data.

One JSON line is printed by rank 0:
  value     whole-job atom-steps/s, state resident in HBM, device time (CUDA events on the integrator's
            stream), max over ranks; re-binning, per-step energy reductions and the per-1000-step speed
            histogram (threadintegrate.cpp:57-62 cadence) are inside the timed region
  e2e       the same metric through the drop-in call pattern of the reference: one integrate() per
            call (host sync each step) that also publishes the step's positions into pinned host buffers
            (ljgpu_step_publish: what get_render_x/y/z expose after every step of the reference; the
            transfer of step k overlaps the computation of step k+1) and an energy read-back per step
  roofline  the dominant kernel (Verlet-list force kernel) against the measured HBM peak, plus its
            FP64 view; cpu_baseline: the reference's own integrator timed on this box's host cores.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "lennard-jones-live-simulator_b200"

DT = 0.001
RHO = 0.8
KT = 1.0


def params_dict(n, rho=RHO, kT=KT):
    return dict(cell_length=float(np.cbrt(n / rho)), nr_particles=int(n), kT=kT, mass=1.0, sigma=1.0,
                rcut=2.5, epsilon=1.0, tau=0.5, shell=0.4, stepsize=DT)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# MEASURED_PEAKS.json has no fp64 entry; measured on this pool's B200 with scripts/fp64_peak.cu (8 independent
# DFMA chains per thread, 64 warps/SM): 33.9 TFLOP/s (nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)
FP64_PEAK_TFLOPS = 33.9


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(n, steps, warm=2):
    """Times the reference's own integrate() (oracle/_ref, else the C restatement) headless on 1 core."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    cfg = params_dict(n)
    if ol.have_ref():
        sim, kind = ol.RefSim(cfg), "reference"
    else:
        sim, kind = ol.Oracle(cfg), "port"
    if warm:
        sim.run(1, warm, DT)
    secs = sim.run(1 + warm, steps, DT)
    rate = n * steps / secs
    sim.close()
    return rate, kind, secs


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU integrator on the host cores, same metric and config."""
    if rank != 0:
        return
    k = max(1, args.steps)
    # bounded sample: the reference runs ~1.1e6 atom-steps/s on one core almost independently of N
    # (BASELINE.md section 2), so the per-step sample is sized for ~120 s of CPU work in total
    budget_atoms = 120.0 * 1.1e6 / (k + args.warmup)
    n = 1048576
    while n > 4096 and n > budget_atoms:
        n //= 2
    t0 = time.time()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    cfg = params_dict(n)
    sim = ol.RefSim(cfg) if ol.have_ref() else ol.Oracle(cfg)
    kind = "reference" if ol.have_ref() else "port"
    if args.warmup:
        sim.run(1, args.warmup, DT)
    secs = sim.run(1 + args.warmup, k, DT)
    rate = n * k / secs
    sample = (f"N={n} atoms at rho*=0.8 (same density/cutoff as the N=1048576 workload, whose CPU rate is "
              f"N-independent to ~40%), start lattice, {k} steps after {args.warmup} warm-up, list rebuilds included; "
              f"1 thread: the reference's force loop is serial and its CMake never enables OpenMP")
    line = {
        "impl": "reference", "metric": "atom-steps/s fp64", "value": rate, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": k, "warmup": args.warmup, "ms_per_step": secs / k * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "N=1048576 LJ liquid rho*=0.8 kT=1 rc=2.5 shell=0.4 tau=0.5 dt=0.001 (BASELINE configs[3])",
                   "sample_atoms": n},
        "cpu_baseline": {"value": rate, "unit": "atom-steps/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count()},
        "e2e": {"value": rate, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def dbg(msg):
    if os.environ.get("LJGPU_BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--atoms", type=int, default=1048576, help="atoms of the whole system (strong scaling over --gpus)")
    ap.add_argument("--atoms-per-gpu", type=int, default=0, help="weak scaling instead: this many atoms per device")
    ap.add_argument("--rho", type=float, default=RHO, help="reduced density (0.84 with --kT 0.72 = triple point, BASELINE configs[4])")
    ap.add_argument("--kT", type=float, default=KT)
    ap.add_argument("--equil", type=int, default=2000, help="untimed equilibration steps before the warm-up")
    ap.add_argument("--kernel", type=int, default=3, help="LJGPU_KERNEL_* (3 = Verlet list)")
    ap.add_argument("--e2e-steps", type=int, default=200)
    ap.add_argument("--profile-every", type=int, default=5,
                    help="CUDA-event stage timing on every k-th step of the timed region (1 = every step, no graph replay)")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the integrator has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local_rank))

    lj = importlib.import_module(PKG)
    weak = args.atoms_per_gpu > 0
    n = args.atoms_per_gpu * world if weak else args.atoms
    cfg = params_dict(n, args.rho, args.kT)
    params = lj.LennardJonesParameters(**cfg)
    lj.host_rng_reset()
    state = lj.host_initialize(params)           # every rank builds the same start state (same generator, same seed)
    if world > 1:
        # z-slab decomposition: one process per device, NCCL halo/migration/all-reduce inside libljgpu
        uid = [lj.slab_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0, device=torch.device("cpu"))
        sim = lj.LennardJonesSimulation(params, state, force_kernel=lj.KERNEL_VERLET, device=local_rank,
                                        slab=(rank, world, uid[0]))
    else:
        sim = lj.LennardJonesSimulation(params, state, force_kernel=args.kernel, device=local_rank)

    dbg(f"created {sim.get_layout()} {sim.slab_info()}")
    step = 1
    sim.integrate_n(step, args.equil, DT); step += args.equil
    dbg("equilibrated")
    sim.integrate_n(step, args.warmup, DT); step += args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: resident state ------------------------------------------------------------
    sim.reset_timers()
    sim.set_profiling(args.profile_every)      # every k-th step is timed kernel by kernel, the rest replay the CUDA graph
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = 0.0
    done = 0
    while done < args.steps:
        chunk = min(1000 - (step % 1000) if step % 1000 else 1000, args.steps - done)
        dev_ms += sim.integrate_n_timed(step, chunk, DT)
        step += chunk; done += chunk
        if step % 1000 == 0:
            sim.get_speed_histogram(100, 10.0)       # signal_velocities cadence, threadintegrate.cpp:57-62
    dbg("timed steps done")
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    timers = sim.get_timers()
    sim.set_profiling(False)
    if world > 1:
        t = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms = float(t.item())
    value = n * args.steps / (dev_ms * 1e-3)
    mean_len, max_len = sim.get_list_stats()
    layout = sim.get_layout()
    dbg(f"value {value:.4e}")

    # ---- timed region 2: end to end through the drop-in call pattern --------------------------------
    # integrate() + publish_state() every step, as the reference does (lennardjonessimulation.cpp:70-98): host call per
    # step (12 bytes of arguments in), energies read back, and this step's positions delivered into one of two pinned
    # host buffer sets (the reference's double buffer); the delivery of step k overlaps the computation of step k+1 and
    # the last one is waited for inside the timed region.
    blocks = [torch.empty(3 * n, dtype=torch.float64, pin_memory=True) for _ in range(2)]     # x | y | z, one block each
    pinned = [[blk[k * n:(k + 1) * n] for k in range(3)] for blk in blocks]
    ptrs = [[p.data_ptr() for p in set_] for set_ in pinned]
    for _ in range(4):
        sim.integrate_publish(step, DT, *ptrs[step & 1]); step += 1
        sim.get_energies()
    sim.publish_wait()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        sim.integrate_publish(step, DT, *ptrs[step & 1]); step += 1
        ek, ep, et = sim.get_energies()
    sim.publish_wait()
    barrier()
    e2e_s = time.perf_counter() - t0
    last = pinned[(step - 1) & 1]
    assert all(bool(torch.isfinite(p).all()) for p in last), "published positions are not finite"
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = n * args.e2e_steps / e2e_s
    dbg(f"e2e {e2e_value:.4e}")
    if world > 1:
        sim.close()
        dist.barrier()

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    hbm_peak, peak_src = measured_peaks()
    # every force launch: plain passes (inner or outer list) + the passes that also prune
    nf = max(1, timers["n_force"] + timers["n_ghost"])
    force_us = (timers["ms_force"] + timers["ms_ghost"]) / nf * 1e3
    n_rc = 65.45 * args.rho                          # ordered in-cutoff neighbours per atom, SURVEY 8(a)/(d)
    if layout["force_kernel"] == 3:
        bytes_per_atom = 24 + 24 + 4 * (mean_len + 1)    # x,y,z in + f out + the atom's list
    else:
        bytes_per_atom = 48
    n_dev = n / world                                 # atoms one launch of the kernel processes (rank 0's device)
    alg_bytes = bytes_per_atom * n_dev
    achieved_gbs = alg_bytes / (force_us * 1e-6) / 1e9
    alg_flops = 28.0 * n_rc * n_dev
    achieved_tf = alg_flops / (force_us * 1e-6) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as fh:
                traffic = json.load(fh).get(f"force_kernel_{layout['force_kernel']}_n{n}")
        except Exception:
            traffic = None
    step_us = dev_ms / args.steps * 1e3
    # "ghost" = force passes that also emitted the pruned inner list (single-device Verlet path)
    stage_share = {k: timers["ms_" + k] / dev_ms for k in ("kick_drift", "force", "ghost", "kick2", "rebuild", "predict", "halo")}

    line = {
        "metric": "atom-steps/s fp64", "value": value, "unit": "atom-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"N={n} LJ liquid rho*={args.rho:g} kT={args.kT:g} rc=2.5 shell=0.4 tau=0.5 dt=0.001 (BASELINE configs[3])",
                   "force_kernel": {1: "allpairs", 2: "cell", 3: "verlet"}[layout["force_kernel"]],
                   "cells_per_axis": layout["cells_per_axis"], "list_capacity": layout["list_capacity"],
                   "mean_list_length": mean_len, "equilibration_steps": args.equil,
                   "l2": "working set (positions, velocities, forces, displacements, list) is several times the 126 MB L2; no flush needed",
                   "parallelism": "single device" if world == 1 else
                   f"{world} z-slabs, NCCL halo exchange + 4-double all-reduce per step, migration at re-binning"},
        "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": 12, "d2h_bytes_per_step": 24 * n + 24,
                "steps": args.e2e_steps, "ms_per_step": e2e_s / args.e2e_steps * 1e3,
                "pattern": "ljgpu_step_publish(step, dt) every step: integrate + positions (3 x f64 x N) delivered into alternating pinned host buffers, transfer of step k overlapping step k+1, last delivery waited for; energies read back after every step"},
        "gpu_launches": int(timers["kernel_launches"]),
        "clocks": clocks,
        "roofline": {"kernel": "tile_kernel<TILE_FORCE_LIST> (Verlet-list force)" if layout["force_kernel"] == 3 else "force kernel",
                     "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_atom": bytes_per_atom, "launch_us": force_us, "launches": int(nf),
                     "share_of_step": stage_share["force"] + stage_share["ghost"],
                     "launches_that_also_pruned": int(timers["n_ghost"])},
        "roofline_fp64": {"kernel": "force kernel", "bound": "fp64", "achieved": achieved_tf, "peak": FP64_PEAK_TFLOPS,
                          "unit": "TFLOP/s", "frac": achieved_tf / FP64_PEAK_TFLOPS,
                          "algorithmic_flops_per_atom": 28.0 * n_rc,
                          "peak_source": "measured DFMA throughput on this pool (scripts/fp64_peak.cu); nominal 37.2"},
        "whole_step": {"us": step_us, "hbm_gbs_at_240B_per_atom_step": 240.0 * n / (step_us * 1e-6) / 1e9 / world,
                       "stage_share": stage_share, "rebuilds": int(timers["rebuilds"]),
                       "stage_us_per_call": {k: timers["ms_" + k] / max(1, timers["n_" + k]) * 1e3
                                             for k in ("kick_drift", "force", "ghost", "kick2", "rebuild", "halo")}},
        "wall_s_timed_region": wall_s,
        "energies_after": {"ekin": ek, "epot": ep, "etot": et},
    }
    if not args.no_cpu and world == 1:
        rate, kind, secs = cpu_reference_rate(n, args.cpu_steps)
        line["cpu_baseline"] = {"value": rate, "unit": "atom-steps/s", "cores": 1, "kind": kind,
                                "host_cores_available": os.cpu_count(),
                                "sample": f"N={n} start lattice, {args.cpu_steps} integrate() calls after 2 warm-up, constructor excluded, {secs:.1f} s"}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
