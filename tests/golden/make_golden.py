#!/usr/bin/env python
"""Writes tests/golden/ref_vectors.json from the reference's OWN integrator (oracle/_ref/libljref.so =
/root/reference/src/simulator/lennardjonessimulation.cpp compiled unmodified, see oracle/Makefile).

Run in the authoring container (where /root/reference exists):   python tests/golden/make_golden.py
Every configuration runs in a fresh process, because the reference's velocity generator is a
function-static that is never reseeded (lennardjonessimulation.cpp:248-252): only the first
simulation of a process sees the default-seeded stream.

All floating-point values are stored as C99 hex strings (float.hex()), i.e. bit-exact.
"""
import json
import os
import subprocess
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as ol  # noqa: E402

DT = 0.001
CONFIGS = {
    "default_n1000": (ol.DEFAULT, 10, True),
    "n4096_rho0.8": (ol.config(4096), 10, True),
    "n4096_triple_point": (ol.config(4096, rho=0.84, kT=0.72), 10, True),
    "n3000_rho0.6_ragged": (ol.config(3000, rho=0.6), 10, True),
    "n65536_rho0.8": (ol.config(65536), 10, False),
    "n1048576_rho0.8": (ol.config(1048576), 3, False),
}


def hx(v):
    return float(v).hex()


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def snapshot(sim):
    x, y, z = sim.positions()
    vx, vy, vz = sim.velocities()
    fx, fy, fz = sim.forces()
    return {
        "crc_positions": [crc(x), crc(y), crc(z)],
        "crc_velocities": [crc(vx), crc(vy), crc(vz)],
        "crc_forces": [crc(fx), crc(fy), crc(fz)],
        "first8": {k: [hx(v) for v in a[:8]] for k, a in
                   dict(x=x, y=y, z=z, vx=vx, vy=vy, vz=vz, fx=fx, fy=fy, fz=fz).items()},
        "sum_v": [hx(vx.sum()), hx(vy.sum()), hx(vz.sum())],
    }


def one(name):
    cfg, steps, with_hist = CONFIGS[name]
    sim = ol.RefSim(cfg)                       # first (and only) simulation of this process
    out = {"config": cfg, "dt": DT, "steps": steps}
    ek, ep, et = sim.energies()
    out["energies"] = [{"step": 0, "ekin": hx(ek), "epot": hx(ep), "etot": hx(et)}]
    # yardstick: the same pair terms accumulated in long double (oracle/lj_oracle.c, ljo_epot_long_double) -
    # on the perfect start lattice the reference's single running double sum is measurably off from it
    out["epot_step0_long_double"] = hx(ol.Oracle(cfg, state=sim.state()).epot_long_double())
    out["list_size_step0"] = sim.list_size()
    cells, dims = sim.cell_ids()
    counts = sim.neighbor_counts()
    incut = sim.incutoff_counts()
    out["step0"] = dict(snapshot(sim), cell_dims=list(dims), crc_cell_ids=crc(cells),
                        crc_neighbor_counts=crc(counts), sum_neighbor_counts=int(counts.sum()),
                        crc_incutoff_counts=crc(incut), sum_incutoff_counts=int(incut.sum()))
    for s in range(1, steps + 1):
        sim.integrate(s, DT)
        ek, ep, et = sim.energies()
        out["energies"].append({"step": s, "ekin": hx(ek), "epot": hx(ep), "etot": hx(et)})
    cells, dims = sim.cell_ids()
    out["final"] = dict(snapshot(sim), crc_cell_ids=crc(cells))
    if with_hist:
        vx, vy, vz = sim.velocities()
        out["final"]["speed_histogram_100_10"] = [int(c) for c in ol.oracle_speed_histogram(vx, vy, vz, 100, 10.0)]
        # the full final state of the small cases travels as a fixture for GPU-side checks
        if cfg["nr_particles"] <= 1000:
            np.savez_compressed(os.path.join(HERE, f"{name}_state_step{steps}.npz"),
                                **dict(zip(("x", "y", "z"), sim.positions())),
                                **dict(zip(("vx", "vy", "vz"), sim.velocities())),
                                **dict(zip(("fx", "fy", "fz"), sim.forces())))
    print(json.dumps({name: out}))


LIQUID_N, LIQUID_MELT, LIQUID_WINDOW, LIQUID_SAMPLE = 1048576, 300, 10, 256


def liquid():
    """BASELINE configs[3] at full size in a LIQUID state, from the reference's own code: 300 integrate() calls melt
    the start lattice (6 list rebuilds), then a 10-step window.  Written to liquid_n1048576.json (energies of steps
    300..310, bit-exact hex) and liquid_n1048576_sample.npz (positions and forces of every 256th atom - 4096 atoms - at
    steps 300 and 310).  The device test starts from the same lattice, integrates the same 300 steps and must land on
    these numbers to ~1e-11: 0.3 time units amplify the last-bit differences of the two force sums only mildly.
    Takes about four minutes of CPU time (the reference needs 0.5-0.7 s per step at this size)."""
    cfg = ol.config(LIQUID_N)
    sim = ol.RefSim(cfg)
    out = {"config": cfg, "dt": DT, "melt_steps": LIQUID_MELT, "window": LIQUID_WINDOW, "sample_every": LIQUID_SAMPLE, "energies": []}
    sim.run(1, LIQUID_MELT, DT)
    arrays = {}

    def sample(tag):
        for name, arr in zip(("x", "y", "z"), sim.positions()):
            arrays[f"{name}_{tag}"] = arr[::LIQUID_SAMPLE].copy()
        for name, arr in zip(("fx", "fy", "fz"), sim.forces()):
            arrays[f"{name}_{tag}"] = arr[::LIQUID_SAMPLE].copy()

    ek, ep, et = sim.energies()
    out["energies"].append({"step": LIQUID_MELT, "ekin": hx(ek), "epot": hx(ep), "etot": hx(et)})
    sample(LIQUID_MELT)
    for s in range(LIQUID_MELT + 1, LIQUID_MELT + LIQUID_WINDOW + 1):
        sim.integrate(s, DT)
        ek, ep, et = sim.energies()
        out["energies"].append({"step": s, "ekin": hx(ek), "epot": hx(ep), "etot": hx(et)})
    sample(LIQUID_MELT + LIQUID_WINDOW)
    cells, dims = sim.cell_ids()
    out["final_crc_cell_ids"] = crc(cells)
    f = np.stack(sim.forces(), axis=1)
    out["force_rms_final"] = hx(np.sqrt((f ** 2).sum(axis=1).mean()))
    out["generator"] = "tests/golden/make_golden.py --liquid (oracle/_ref/libljref.so = the reference's sources compiled unmodified)"
    with open(os.path.join(HERE, "liquid_n1048576.json"), "w") as fh:
        json.dump(out, fh, indent=1)
    np.savez_compressed(os.path.join(HERE, "liquid_n1048576_sample.npz"), **arrays)
    print("wrote liquid_n1048576.json / liquid_n1048576_sample.npz", file=sys.stderr)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--liquid":
        liquid()
        return
    if len(sys.argv) > 1:
        one(sys.argv[1])
        return
    if not ol.have_ref():
        raise SystemExit("oracle/_ref/libljref.so missing: run `make -C oracle` where /root/reference exists")
    allv = {"generator": "tests/golden/make_golden.py", "source": "oracle/_ref/libljref.so (reference sources compiled unmodified, g++ -O3, no -march, no OpenMP)"}
    for name in CONFIGS:
        res = subprocess.run([sys.executable, os.path.abspath(__file__), name], check=True, capture_output=True, text=True)
        allv.update(json.loads(res.stdout.strip().splitlines()[-1]))
        print("done", name, file=sys.stderr)
    with open(os.path.join(HERE, "ref_vectors.json"), "w") as fh:
        json.dump(allv, fh, indent=1)


if __name__ == "__main__":
    main()
