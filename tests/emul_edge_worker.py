"""Worker of tests/test_emulated_library.py::test_edge_cases_on_the_emulated_library: small, odd and extreme
configurations through the emulated product library (LJGPU_LIBRARY) against the oracle.  Test infrastructure only."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as ol  # noqa: E402

lj = importlib.import_module("lennard-jones-live-simulator_b200")
BASE = dict(kT=1.0, mass=1.0, sigma=1.0, rcut=2.5, epsilon=1.0, tau=0.5, shell=0.4, stepsize=0.001)


def cfg(n, L, **kw):
    d = dict(BASE, cell_length=L, nr_particles=n)
    d.update(kw)
    return d


def run(tag, c, kernel, steps=12, dt=0.001, tol=1e-12, min_rebuilds=0):
    o = ol.Oracle(c)
    g = lj.LennardJonesSimulation(lj.LennardJonesParameters(**c), o.state(), force_kernel=kernel)
    for s in range(0, steps + 1):
        if s:
            o.integrate(s, dt); g.integrate(s, dt)
        err = ol.force_rel_error(g.get_forces_soa(), o.forces(), o.force_abs_sums()).max()
        ek, ep, _ = g.get_energies()
        oek, oep, _ = o.energies()
        own = abs(oep - o.epot_long_double())
        assert err < tol, (tag, kernel, s, "force", err)
        assert abs(ep - oep) <= tol * abs(oep) + 2 * own, (tag, kernel, s, "epot", ep, oep)
        if oek:
            assert abs(ek - oek) <= tol * abs(oek), (tag, kernel, s, "ekin", ek, oek)
    cells, _ = g.get_cell_ids()
    x, y, z = g.get_positions_soa()
    assert np.array_equal(cells, ol.oracle_cell_ids_of(c, x, y, z)[0]), (tag, kernel, "cell ids")
    nb = g.get_neighbor_counts(c["rcut"] + c["shell"], True)
    assert int(nb.sum()) == int(ol.Oracle(c, state=(x, y, z) + tuple(g.get_velocities_soa())).neighbor_counts().sum()), (tag, "counts")
    assert g.get_timers()["rebuilds"] >= min_rebuilds, (tag, kernel, g.get_timers()["rebuilds"])
    g.close()
    print(f"ok {tag} kernel {kernel}")


CASES = [
    ("two atoms", cfg(2, 9.0), (1, 2, 3), {}),
    ("three cells per axis, dense", cfg(700, 8.88), (2, 3), {}),
    ("dilute, most cells empty", cfg(2000, 34.2), (2, 3), {}),
    ("sigma 1.2, epsilon 0.7, mass 2, rcut 3", cfg(1500, 16.0, sigma=1.2, epsilon=0.7, mass=2.0, rcut=3.0, shell=0.5), (1, 2, 3), {}),
    ("tiny shell: re-binning every few steps", cfg(2048, 14.0, shell=0.05), (2, 3), dict(min_rebuilds=3)),
    ("wide shell", cfg(4096, 17.2355, shell=1.0), (3,), {}),
    ("negative tau", cfg(1000, 14.938, tau=-0.5), (3,), {}),
    ("dt = 0.004 for 60 steps", cfg(4096, 17.2355), (3,), dict(steps=60, dt=0.004, tol=1e-9, min_rebuilds=4)),
]

if __name__ == "__main__":
    for tag, c, kernels, kw in CASES:
        for k in kernels:
            run(tag, c, k, **kw)
    print("EDGE CASES OK")
