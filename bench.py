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
  parity    energies of steps 1-3 from the reference's start state against the golden vectors written by the
            reference's own code (tests/golden/ref_vectors.json), at every N - so the multi-GPU lines carry a
            parity check of the slab path too
  roofline  the dominant kernel (the Verlet-list force pass) against the FP64 peak measured in this run
            (ljgpu_measure_fp64_peak; SURVEY.md 8d: the force kernels are FP64-bound, 28 n_rc flop per atom), its HBM
            view beside it; kernel times come from a dedicated window in which EVERY step is timed stage by stage
  configs   the other single-device BASELINE configurations in the same invocation (N = 4096 all-pairs,
            N = 65 536 cell list and Verlet list), each with its own kernel roofline; with 8 GPUs also
            BASELINE configs[4] (N = 16.7 M at the triple point, weak scaling)
  cpu_baseline  the reference's own integrator timed on this box's host cores.
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


FP64_NOMINAL_TFLOPS = 37.2      # 148 SMs x 64 DFMA/clk x 2 x 1.965 GHz; the denominator actually used is measured in the run
WORKLOAD = "N=1048576 LJ liquid rho*=0.8 kT=1 rc=2.5 shell=0.4 tau=0.5 dt=0.001 (BASELINE configs[3])"


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


def workload_string(n, rho, kT):
    if n == 1048576 and rho == RHO and kT == KT:
        return WORKLOAD
    return f"N={n} LJ liquid rho*={rho:g} kT={kT:g} rc=2.5 shell=0.4 tau=0.5 dt=0.001"


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU integrator on the host cores, same metric, same config (the whole
    N = 2^20 system, the reference's own start state), 1 thread - its force loop is serial and its CMake never enables
    OpenMP.  A step costs ~0.4-0.7 s, so K + W steps stay within a minute or two for the K the driver uses; for a large
    K the timed steps are capped at 120 and the rate reported is that of the steps actually run."""
    if rank != 0:
        return
    weak = args.atoms_per_gpu > 0
    n = args.atoms_per_gpu * world if weak else args.atoms
    k = max(1, min(args.steps, 120))
    w = min(args.warmup, 10)
    t0 = time.time()
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    cfg = params_dict(n, args.rho, args.kT)
    sim = ol.RefSim(cfg) if ol.have_ref() else ol.Oracle(cfg)
    kind = "reference" if ol.have_ref() else "port"
    if w:
        sim.run(1, w, DT)
    secs = sim.run(1 + w, k, DT)
    rate = n * k / secs
    sample = (f"the whole workload (N={n}), reference's start lattice, {k} integrate() calls after {w} warm-up calls, constructor "
              f"(lattice, first list, first forces) excluded, {secs:.1f} s; 1 thread: the reference's force loop is serial and its "
              f"CMake never enables OpenMP; no list rebuild falls into so short a window on the lattice - with rebuilds in the "
              f"window the reference runs at about half this rate (BASELINE.md)")
    line = {
        "impl": "reference", "metric": "atom-steps/s fp64", "value": rate, "unit": "atom-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / k * 1e3,
        "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(n, args.rho, args.kT)},
        "cpu_baseline": {"value": rate, "unit": "atom-steps/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores_available": os.cpu_count(), "steps_timed": k},
        "e2e": {"value": rate, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.time() - t0,
    }
    print(json.dumps(line), flush=True)


def dbg(msg):
    if os.environ.get("LJGPU_BENCH_DEBUG"):
        print(f"[bench rank {os.environ.get('RANK', '0')}] {msg}", file=sys.stderr, flush=True)


def golden_energies(n, rho, kT):
    """Energies of steps 1-3 written by the reference's own code (tests/golden/make_golden.py), or None."""
    if not (n == 1048576 and rho == RHO and kT == KT):
        return None
    try:
        with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")) as fh:
            g = json.load(fh)["n1048576_rho0.8"]
        return {r["step"]: (float.fromhex(r["ekin"]), float.fromhex(r["epot"])) for r in g["energies"] if r["step"] > 0}
    except Exception:
        return None


def stage_window(sim, first_step, nsteps):
    """A dedicated window in which EVERY step is timed kernel by kernel (plain launches, CUDA events around each stage on
    the integrator's stream).  Returns per-stage {us_per_call, calls, share of the window's device time} and the
    window's device time per step."""
    sim.reset_timers()
    sim.set_profiling(1)
    dev_ms = sim.integrate_n_timed(first_step, nsteps, DT)
    t = sim.get_timers()
    sim.set_profiling(False)
    stages = {}
    for k in ("kick_drift", "force", "force_prune", "kick2", "rebuild", "predict", "halo"):
        calls = int(t["n_" + k])
        stages[k] = {"us_per_call": t["ms_" + k] / calls * 1e3 if calls else None, "calls": calls,
                     "share": t["ms_" + k] / dev_ms if dev_ms > 0 else None}
    return stages, dev_ms / nsteps * 1e3, t


def force_roofline(stages, n_dev, rho, kernel_name, fp64_peak, hbm_peak, peak_src, mean_len=None, traffic=None, flops_per_atom=None):
    """FP64 roofline of the force kernel: algorithmic flops = 28 flop per ordered in-cutoff pair,
    n_rc = 65.45 rho* pairs per atom (all-pairs kernel: + 11 flop per rejected candidate, the path is O(N^2) by
    specification), divided by the average launch duration of ALL force launches of the window (plain passes and the
    passes that also prune)."""
    calls = stages["force"]["calls"] + stages["force_prune"]["calls"]
    if calls == 0:
        return None
    us = ((stages["force"]["us_per_call"] or 0.0) * stages["force"]["calls"] +
          (stages["force_prune"]["us_per_call"] or 0.0) * stages["force_prune"]["calls"]) / calls
    n_rc = 65.45 * rho
    fpa = flops_per_atom if flops_per_atom is not None else 28.0 * n_rc
    achieved = fpa * n_dev / (us * 1e-6) / 1e12
    hbm_bytes = 48.0 * n_dev                    # compulsory: x,y,z in + f out
    return {"kernel": kernel_name, "bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": achieved / fp64_peak, "frac_of_nominal": achieved / FP64_NOMINAL_TFLOPS,
            "peak_source": "ljgpu_measure_fp64_peak in this run (8 independent DFMA chains x 64 warps/SM); nominal 37.2",
            "algorithmic_flops_per_atom": fpa, "atoms_per_launch": n_dev, "launch_us": us, "launches": calls,
            "launches_that_also_pruned": stages["force_prune"]["calls"],
            "share_of_step": (stages["force"]["share"] or 0.0) + (stages["force_prune"]["share"] or 0.0),
            "traffic": traffic,
            "hbm_view": {"algorithmic_bytes_per_atom": 48, "achieved_gbs": hbm_bytes / (us * 1e-6) / 1e9, "peak_gbs": hbm_peak,
                         "frac": hbm_bytes / (us * 1e-6) / 1e9 / hbm_peak, "peak_source": peak_src,
                         "reference_list_format_bytes_per_atom": (48 + 4 * (mean_len + 1)) if mean_len else None,
                         "measured_dram_bytes_per_launch": traffic}}


def small_config_block(lj, n, kernel, device, fp64_peak, hbm_peak, peak_src, equil, steps):
    """One of the other single-device BASELINE configurations, measured the same way (resident state, device time)."""
    cfg = params_dict(n)
    params = lj.LennardJonesParameters(**cfg)
    lj.host_rng_reset()
    sim = lj.LennardJonesSimulation(params, lj.host_initialize(params), force_kernel=kernel, device=device)
    sim.integrate_n(1, equil, DT)
    step = equil + 1
    sim.integrate_n(step, 50, DT); step += 50
    dev_ms = sim.integrate_n_timed(step, steps, DT); step += steps
    stages, _, t = stage_window(sim, step, min(steps, 300))
    name = {1: "force_allpairs_kernel (all j through shared memory)", 2: "tile_kernel<TILE_FORCE_CELL> (27-cell stencil)",
            3: "tile_kernel<TILE_FORCE_LIST> (Verlet list)"}[kernel]
    fpa = 11.0 * (n - 1) + 17.0 * 65.45 * RHO if kernel == 1 else None
    ek, ep, _ = sim.get_energies()
    out = {"workload": f"N={n} LJ liquid rho*=0.8 kT=1 rc=2.5 ({'BASELINE configs[1]' if n == 4096 else 'BASELINE configs[2]'})",
           "force_kernel": {1: "allpairs", 2: "cell", 3: "verlet"}[kernel], "value": n * steps / (dev_ms * 1e-3), "unit": "atom-steps/s",
           "ms_per_step": dev_ms / steps, "steps": steps, "equilibration_steps": equil,
           "roofline": force_roofline(stages, n, RHO, name, fp64_peak, hbm_peak, peak_src, flops_per_atom=fpa),
           "stage_us_per_call": {k: v["us_per_call"] for k, v in stages.items() if v["calls"]},
           "energy_per_atom": {"ekin": ek / n, "epot": ep / n}}
    sim.close()
    return out


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
    ap.add_argument("--stage-steps", type=int, default=200, help="length of the dedicated window in which every step is timed stage by stage")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the blocks for the other BASELINE configurations")
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
    t_bench0 = time.time()
    fp64_peak = lj.measure_fp64_peak(local_rank)      # the force kernels' roofline denominator, measured here
    hbm_peak, peak_src = measured_peaks()

    def make_sim(n, rho, kT, kernel):
        cfg = params_dict(n, rho, kT)
        params = lj.LennardJonesParameters(**cfg)
        lj.host_rng_reset()
        state = lj.host_initialize(params)       # every rank builds the same start state (same generator, same seed)
        if world > 1:
            # z-slab decomposition: one process per device; peer-memory halo push + mailbox all-reduce per step inside libljgpu
            uid = [lj.slab_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0, device=torch.device("cpu"))
            return lj.LennardJonesSimulation(params, state, force_kernel=lj.KERNEL_VERLET, device=local_rank, slab=(rank, world, uid[0]))
        return lj.LennardJonesSimulation(params, state, force_kernel=kernel, device=local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    weak = args.atoms_per_gpu > 0
    n = args.atoms_per_gpu * world if weak else args.atoms
    sim = make_sim(n, args.rho, args.kT, args.kernel)
    dbg(f"created {sim.get_layout()} {sim.slab_info()}")

    # ---- parity: steps 1-3 from the reference's start state against the reference's own golden energies -----------
    gold = golden_energies(n, args.rho, args.kT)
    parity = {"ok": None, "reason": "no golden vectors for this workload"}
    step = 1
    if gold:
        worst = 0.0
        for s_ in (1, 2, 3):
            sim.integrate(s_, DT)
            ek_, ep_, _ = sim.get_energies()
            worst = max(worst, abs(ek_ - gold[s_][0]) / abs(gold[s_][0]), abs(ep_ - gold[s_][1]) / abs(gold[s_][1]))
        step = 4
        parity = {"max_rel_err": worst, "ok": bool(worst < 1e-11), "tolerance": 1e-11,
                  "what": "ekin and epot of integrate() steps 1-3 from the reference's start state vs tests/golden/ref_vectors.json "
                          "(written by the reference's own code); 1e-11 because the reference's single running sum over 5.5e7 pair terms "
                          "is itself ~6e-10 off on the lattice at step 0 and ~1e-12 off in the first steps"}
    sim.integrate_n(step, args.equil, DT); step += args.equil
    dbg("equilibrated")
    sim.integrate_n(step, args.warmup, DT); step += args.warmup

    # ---- timed region 1: resident state (CUDA-graph replay on one device), exactly --steps steps --------------------
    sim.reset_timers()
    sim.set_profiling(False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)                            # let the sampler deliver its first lines: the region itself may last milliseconds
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
    barrier()
    wall_s = time.perf_counter() - t_wall0
    timers_value = sim.get_timers()
    dev_ms = max_over_ranks(dev_ms)
    value = n * args.steps / (dev_ms * 1e-3)
    dbg(f"value {value:.4e}")

    # ---- the same state, a dedicated window with every step timed stage by stage (feeds `roofline`) ------------------
    stages, window_us_per_step, t_stage = stage_window(sim, step, args.stage_steps)
    step += args.stage_steps
    clocks = sampler.stop()
    mean_len, max_len = sim.get_list_stats()
    layout = sim.get_layout()

    # ---- timed region 2: end to end through the drop-in call pattern --------------------------------
    # integrate() + publish_state() every step, as the reference does (lennardjonessimulation.cpp:70-98): host call per
    # step (12 bytes of arguments in), energies read back, and this step's positions delivered into pinned host memory
    # (two buffer sets, the reference's double buffer); the delivery of step k overlaps the computation of step k+1 and
    # the last one is waited for inside the timed region.  On slabs every device delivers the atoms it owns straight
    # into ONE frame shared by all ranks (POSIX shared memory registered with CUDA), 24 N bytes over PCIe per step in total.
    shm = lj.SharedFrames(n, world, rank, dist if world > 1 else None)
    ptrs = shm.pointers()
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
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert shm.frame_is_finite((step - 1) & 1), "published positions are not finite"
    e2e_value = n * args.e2e_steps / e2e_s
    dbg(f"e2e {e2e_value:.4e}")
    slab_info = sim.slab_info()
    sim.close()
    shm.close()
    barrier()

    # ---- the other BASELINE configurations -------------------------------------------------------------------------
    blocks = {}
    if not args.no_configs and not weak and n == 1048576:
        if world == 1:
            blocks["n4096_allpairs"] = small_config_block(lj, 4096, 1, local_rank, fp64_peak, hbm_peak, peak_src, 2000, 2000)
            blocks["n65536_cell"] = small_config_block(lj, 65536, 2, local_rank, fp64_peak, hbm_peak, peak_src, 1500, 300)
            blocks["n65536_verlet"] = small_config_block(lj, 65536, 3, local_rank, fp64_peak, hbm_peak, peak_src, 1500, 1000)
        if world == 8:
            # BASELINE configs[4]: N = 16 777 216 at the triple point, 2 097 152 atoms per device (weak scaling)
            nw = 2097152 * world
            simw = make_sim(nw, 0.84, 0.72, 3)
            sw = 1
            simw.integrate_n(sw, 1000, DT); sw += 1000
            barrier()
            ms_w = max_over_ranks(simw.integrate_n_timed(sw, 300, DT)); sw += 300
            stages_w, _, _ = stage_window(simw, sw, 100)
            ekw, epw, _ = simw.get_energies()
            blocks["weak_16M"] = {"workload": "N=16777216 LJ fluid at the triple point rho*=0.84 kT=0.72, 8 z-slabs of 2097152 atoms (BASELINE configs[4])",
                                  "value": nw * 300 / (ms_w * 1e-3), "unit": "atom-steps/s", "ms_per_step": ms_w / 300, "steps": 300,
                                  "equilibration_steps": 1000, "target": 1e10,
                                  "roofline": force_roofline(stages_w, nw / world, 0.84, "tile_kernel<TILE_FORCE_LIST> (Verlet list), rank 0's slab",
                                                             fp64_peak, hbm_peak, peak_src),
                                  "energy_per_atom": {"ekin": ekw / nw, "epot": epw / nw}}
            simw.close()
            barrier()

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- roofline of the dominant kernel -----------------------------------------------------------
    n_dev = n / world                                 # atoms one launch of the kernel processes (rank 0's device)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            with open(tpath) as fh:
                traffic = json.load(fh).get(f"force_kernel_{layout['force_kernel']}_atoms_per_device_{int(n_dev)}_rho{args.rho:g}")
        except Exception:
            traffic = None
    kernel_name = {1: "force_allpairs_kernel", 2: "tile_kernel<TILE_FORCE_CELL>", 3: "tile_kernel<TILE_FORCE_LIST / TILE_FORCE_PRUNE> (Verlet-list force pass)"}[layout["force_kernel"]]
    roofline = force_roofline(stages, n_dev, args.rho, kernel_name, fp64_peak, hbm_peak, peak_src, mean_len=mean_len, traffic=traffic)
    step_us = dev_ms / args.steps * 1e3
    n_rc = 65.45 * args.rho

    line = {
        "metric": "atom-steps/s fp64", "value": value, "unit": "atom-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_string(n, args.rho, args.kT)},
        "setup": {"force_kernel": {1: "allpairs", 2: "cell", 3: "verlet"}[layout["force_kernel"]],
                  "cells_per_axis": layout["cells_per_axis"], "list_capacity": layout["list_capacity"],
                  "mean_list_length": mean_len, "equilibration_steps": args.equil,
                  "state": "reference start lattice + Gaussian velocities, steps 1-3 checked against the reference (parity), then equilibrated",
                  "l2": "working set (positions, velocities, forces, displacements, list) is several times the 126 MB L2; no flush needed",
                  "parallelism": "single device" if world == 1 else
                  f"{world} z-slabs, one process per GPU: peer-memory (NVLink) halo push + mailbox all-reduce of 4 doubles per step, NCCL only for migration at re-binning",
                  "slab_rank0": slab_info if world > 1 else None},
        "e2e": {"value": e2e_value, "unit": "atom-steps/s", "h2d_bytes_per_step": 12 * world, "d2h_bytes_per_step": 24 * n + 24 * world,
                "steps": args.e2e_steps, "ms_per_step": e2e_s / args.e2e_steps * 1e3,
                "pattern": "ljgpu_step_publish(step, dt) every step: integrate + positions (3 x f64 x N) delivered into alternating pinned host frames, "
                           "transfer of step k overlapping step k+1, last delivery waited for; energies read back after every step" +
                           ("" if world == 1 else "; every device delivers the atoms it owns into one frame shared by all ranks")},
        "gpu_launches": int(timers_value["kernel_launches"]),
        "clocks": clocks,
        "parity": parity,
        "roofline": roofline,
        "whole_step": {"us": step_us,
                       "fp64_tflops_at_28nrc_plus_56_flop": (28.0 * n_rc + 56.0) * n / (step_us * 1e-6) / 1e12 / world,
                       "hbm_gbs_at_240B_per_atom_step": 240.0 * n / (step_us * 1e-6) / 1e9 / world,
                       "rebuilds_in_timed_region": int(timers_value["rebuilds"]),
                       "stage_window": {"steps": args.stage_steps, "us_per_step": window_us_per_step, "rebuilds": int(t_stage["rebuilds"]),
                                        "how": "every step of this window is timed stage by stage with CUDA events on the integrator's stream (plain launches; "
                                               "the timed region above replays CUDA graphs); shares are fractions of the window's own device time",
                                        "stages": stages}},
        "wall_s_timed_region": wall_s,
        "energies_after": {"ekin": ek, "epot": ep, "etot": et},
        "configs": blocks,
    }
    if not args.no_cpu and world == 1:
        rate, kind, secs = cpu_reference_rate(n, args.cpu_steps)
        line["cpu_baseline"] = {"value": rate, "unit": "atom-steps/s", "cores": 1, "kind": kind,
                                "host_cores_available": os.cpu_count(),
                                "sample": f"N={n} start lattice, {args.cpu_steps} integrate() calls after 2 warm-up, constructor excluded, {secs:.1f} s"}
    line["wall_s_bench"] = time.time() - t_bench0
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
