"""Developer probe: time the integrator at a given N / kernel with per-stage CUDA-event timers."""
import importlib, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
lj = importlib.import_module("lennard-jones-live-simulator_b200")

def run(n, kernel, steps, rho=0.8, kT=1.0, warm=200, dt=0.001):
    L = float(np.cbrt(n / rho))
    p = lj.LennardJonesParameters(cell_length=L, nr_particles=n, kT=kT, mass=1.0, sigma=1.0, rcut=2.5,
                                  epsilon=1.0, tau=0.5, shell=0.4, stepsize=0.001)
    lj.host_rng_reset()
    t0 = time.time(); st = lj.host_initialize(p); t1 = time.time()
    g = lj.LennardJonesSimulation(p, st, force_kernel=kernel); t2 = time.time()
    g.integrate_n(1, warm, 0.001)
    # dt = 0 (timing experiments whose forces are wrong on purpose): the atoms stay where the warm-up left them
    g.reset_timers(); g.set_profiling(True)
    t3 = time.time(); g.integrate_n(warm + 1, steps, dt); e = g.get_energies(); t4 = time.time()
    tm = g.get_timers(); g.set_profiling(False)
    t5 = time.time(); g.integrate_n(warm + steps + 1, steps, dt); e = g.get_energies(); t6 = time.time()
    print(f"N={n} kernel={kernel} layout={g.get_layout()} init {t1-t0:.2f}s create {t2-t1:.2f}s")
    print(f"  profiled: {n*steps/(t4-t3):.3e} atom-steps/s   unprofiled: {n*steps/(t6-t5):.3e} atom-steps/s  ({(t6-t5)/steps*1e6:.1f} us/step)")
    for k in ("predict", "kick_drift", "force_prune", "force", "kick2", "rebuild"):
        ms, c = tm["ms_" + k], tm["n_" + k]
        if c: print(f"  {k:10s} {ms:9.3f} ms total  {c:6d} calls  {ms/c*1e3:9.2f} us/call   {ms/ (t4-t3)/10:.1f}% of wall")
    print("  energies", e, "rebuilds", tm["rebuilds"], "launches", tm["kernel_launches"])
    g.close()

if __name__ == "__main__":
    n = int(sys.argv[1]); kernel = int(sys.argv[2]); steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
    warm = int(sys.argv[4]) if len(sys.argv) > 4 else 200
    dt = float(sys.argv[5]) if len(sys.argv) > 5 else 0.001
    run(n, kernel, steps, warm=warm, dt=dt)
