"""Kernel logic without a GPU: the product's device headers (lj_tile.cuh) compiled by g++ against
tests/cuda_emul/cuda_emul.h (threads of a block as host threads) and run on one configuration - tile tables,
list count + fill, force pass over the outer list, force pass that also prunes, force pass over the pruned list -
checked against the oracle.  Test infrastructure only: nothing here is linked into libljgpu.so, the product has no
CPU path, and the parity tests proper are the `-m gpu` ones."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "cuda_emul")
DT = 0.001


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL, "libljemul.so")
    src = os.path.join(EMUL, "emul_kernels.cpp")
    csrc = os.path.join(os.path.dirname(HERE), "lennard-jones-live-simulator_b200", "csrc")
    deps = [src, os.path.join(EMUL, "cuda_emul.h")] + [os.path.join(csrc, f) for f in
                                                        ("lj_common.cuh", "lj_kernels.cuh", "lj_tile.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-mfma", "-ffp-contract=off",
                        "-DLJ_EMUL", "-o", so, src], check=True, cwd=EMUL, timeout=300)
    lib = C.CDLL(so)
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    lib.ljemul_forces.argtypes = [C.c_uint, C.c_uint] + [C.c_double] * 6 + [C.c_uint] * 3 + [dp] * 5 + [up, up]
    lib.ljemul_forces.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def liquid():
    """N = 10 000 at rho* = 0.8 (8 cells per axis), 150 steps off the lattice; forces and list lengths of that
    state from a fresh oracle."""
    cfg = ol.config(10000)
    o = ol.Oracle(cfg)
    o.run(1, 150, DT)
    ref = ol.Oracle(cfg, state=o.state())          # constructor: list + forces of exactly these positions
    return cfg, ref


@pytest.mark.parametrize("ctas,brick", [(1000, (3, 2, 2)), (5, (3, 2, 2)), (7, (1, 1, 1)), (3, (2, 2, 1))],
                         ids=["3x2x2-one-brick-per-cta", "3x2x2-streamed", "1x1x1-streamed", "2x2x1-streamed"])
def test_emulated_kernels_match_oracle(emul, liquid, ctas, brick):
    """ctas < bricks: every CTA streams several bricks through its two-slot ring (staging warp / computing warps,
    full and empty barriers), which is how the product runs at N >= 65 536."""
    cfg, ref = liquid
    n, L = cfg["nr_particles"], cfg["cell_length"]
    x, y, z = [np.ascontiguousarray(a) for a in ref.positions()]
    dp, up = C.POINTER(C.c_double), C.POINTER(C.c_uint32)
    F = np.zeros((4, 3, n)); E = np.zeros(4)
    counts = np.zeros(n, dtype=np.uint32); info = np.zeros(8, dtype=np.uint32)
    rc = emul.ljemul_forces(ctas, n, L, cfg["rcut"], cfg["shell"], cfg["sigma"], cfg["epsilon"], 0.12, *brick,
                            x.ctypes.data_as(dp), y.ctypes.data_as(dp), z.ctypes.data_as(dp),
                            F.ctypes.data_as(dp), E.ctypes.data_as(dp), counts.ctypes.data_as(up), info.ctypes.data_as(up))
    assert rc == 0, (rc, info)
    # list lengths: bit-exact (integer output); forces: 1e-12; the three passes: bit-identical
    assert np.array_equal(counts, ref.neighbor_counts())
    err = ol.force_rel_error(tuple(F[0]), ref.forces(), ref.force_abs_sums()).max()
    assert err < 1e-12, err
    assert np.array_equal(F[0], F[1]) and np.array_equal(F[0], F[2])
    assert E[0] == E[1] == E[2]
    _, epot, _ = ref.energies()
    exact = ref.epot_long_double()
    assert abs(E[0] * 4.0 * cfg["epsilon"] * 0.5 - exact) < 1e-12 * abs(exact) + 2 * abs(epot - exact)
    assert info[6] == 1 and 0 < info[7]            # the third pass did walk the pruned list
    # the cell-list force kernel (no list): same forces to 1e-12, same energy to rounding
    err_cell = ol.force_rel_error(tuple(F[3]), ref.forces(), ref.force_abs_sums()).max()
    assert err_cell < 1e-12, err_cell
    assert abs(E[3] - E[0]) < 1e-13 * abs(E[0])
