"""CPU tests: the C restatement (oracle/lj_oracle.c) is pinned against
  (1) tests/golden/ref_vectors.json, written by tests/golden/make_golden.py from the reference's own
      sources compiled unmodified (oracle/_ref), and
  (2) oracle/_ref itself, bit for bit, whenever that library is present (authoring container and,
      as a prebuilt file, the GPU box)."""
import json
import os
import zlib

import numpy as np
import pytest

import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
DT = 0.001


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "ref_vectors.json")) as fh:
        return json.load(fh)


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def fh(s):
    return float.fromhex(s)


SMALL = ["default_n1000", "n4096_rho0.8", "n4096_triple_point", "n3000_rho0.6_ragged", "n65536_rho0.8"]


@pytest.mark.parametrize("name", SMALL)
def test_oracle_reproduces_reference_golden_vectors(golden, name):
    g = golden[name]
    cfg = g["config"]
    o = ol.Oracle(cfg)                                   # own generator restarted: fresh-process stream
    assert o.list_size() == g["list_size_step0"]
    cells, dims = o.cell_ids()
    assert list(dims) == g["step0"]["cell_dims"]
    assert crc(cells) == g["step0"]["crc_cell_ids"]
    counts = o.neighbor_counts()
    assert crc(counts) == g["step0"]["crc_neighbor_counts"] and int(counts.sum()) == g["step0"]["sum_neighbor_counts"]
    incut = o.incutoff_counts()
    assert crc(incut) == g["step0"]["crc_incutoff_counts"] and int(incut.sum()) == g["step0"]["sum_incutoff_counts"]
    assert [crc(a) for a in o.positions()] == g["step0"]["crc_positions"]
    assert [crc(a) for a in o.velocities()] == g["step0"]["crc_velocities"]
    assert [crc(a) for a in o.forces()] == g["step0"]["crc_forces"]
    for rec in g["energies"]:
        if rec["step"] > 0:
            o.integrate(rec["step"], DT)
        ek, ep, et = o.energies()
        assert (ek, ep, et) == (fh(rec["ekin"]), fh(rec["epot"]), fh(rec["etot"])), f"{name} step {rec['step']}"
    fin = g["final"]
    assert [crc(a) for a in o.positions()] == fin["crc_positions"]
    assert [crc(a) for a in o.velocities()] == fin["crc_velocities"]
    assert [crc(a) for a in o.forces()] == fin["crc_forces"]
    assert crc(o.cell_ids()[0]) == fin["crc_cell_ids"]
    x = o.positions()[0]
    assert [float(v).hex() for v in x[:8]] == fin["first8"]["x"]
    if "speed_histogram_100_10" in fin:
        vx, vy, vz = o.velocities()
        assert [int(c) for c in ol.oracle_speed_histogram(vx, vy, vz)] == fin["speed_histogram_100_10"]


def test_known_answers_of_the_default_configuration(golden):
    """The values SURVEY.md 8(c) lists for the shipped default configuration (mainwindow.cpp:67-77)."""
    e = {r["step"]: r for r in golden["default_n1000"]["energies"]}
    assert fh(e[0]["ekin"]) == 0.0                        # ekin reads 0 before the first step
    assert fh(e[0]["epot"]) == -1102.9177850342294
    assert fh(e[1]["ekin"]) == 1474.4462944100519 and fh(e[1]["epot"]) == -1102.9276615686763
    assert fh(e[10]["ekin"]) == 1475.8474496253316 and fh(e[10]["etot"]) == 371.9418271604909
    assert golden["default_n1000"]["list_size_step0"] == 28000
    assert fh(golden["n1048576_rho0.8"]["energies"][0]["epot"]) == -5017788.9669708563
    assert golden["n1048576_rho0.8"]["list_size_step0"] == 85316638


def test_golden_state_fixture_matches_oracle(golden):
    d = np.load(os.path.join(HERE, "golden", "default_n1000_state_step10.npz"))
    o = ol.Oracle(ol.DEFAULT)
    o.run(1, 10, DT)
    for k, a in zip(("x", "y", "z"), o.positions()):
        assert np.array_equal(d[k], a)
    for k, a in zip(("vx", "vy", "vz"), o.velocities()):
        assert np.array_equal(d[k], a)
    for k, a in zip(("fx", "fy", "fz"), o.forces()):
        assert np.array_equal(d[k], a)


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libljref.so not built (needs /root/reference)")
def test_oracle_bit_identical_to_compiled_reference_across_rebuilds():
    """300 steps of the default configuration incl. ~6 list rebuilds; state injected so the comparison
    does not depend on which simulation the reference's static generator served first."""
    cfg = ol.DEFAULT
    o = ol.Oracle(cfg)
    r = ol.RefSim(cfg, state=o.state())
    assert o.energies() == r.energies()
    for s in range(1, 301):
        o.integrate(s, DT)
        r.integrate(s, DT)
        assert o.energies() == r.energies(), s
    for a, b in zip(o.positions() + o.velocities() + o.forces() + o.dij(),
                    r.positions() + r.velocities() + r.forces() + r.dij()):
        assert np.array_equal(a, b)
    assert o.rebuild_count() >= 5
    assert o.list_size() == r.list_size()
    assert np.array_equal(o.neighbor_counts(), r.neighbor_counts())
    assert np.array_equal(o.incutoff_counts(), r.incutoff_counts())
    assert np.array_equal(o.cell_ids()[0], r.cell_ids()[0])


@pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libljref.so not built (needs /root/reference)")
def test_oracle_bit_identical_to_compiled_reference_dense_liquid():
    cfg = ol.config(4096, rho=0.84, kT=0.72, tau=1e300)        # NVE through the unchanged API
    o = ol.Oracle(cfg)
    r = ol.RefSim(cfg, state=o.state())
    for s in range(1, 61):
        o.integrate(s, DT)
        r.integrate(s, DT)
    assert o.energies() == r.energies()
    for a, b in zip(o.positions() + o.velocities(), r.positions() + r.velocities()):
        assert np.array_equal(a, b)


def test_oracle_movie_writer_format(tmp_path):
    """u32 N on create, then per frame 3 f64 box + N x 7 f64 (lennardjonessimulation.cpp:121-157)."""
    o = ol.Oracle(ol.DEFAULT)
    path = str(tmp_path / "movie.bin")
    o.write_movie(path, True)
    o.integrate(1, DT)
    o.write_movie(path, False)
    raw = open(path, "rb").read()
    n = 1000
    assert len(raw) == 4 + 2 * (24 + n * 56)
    assert int(np.frombuffer(raw[:4], dtype=np.uint32)[0]) == n
    f2 = np.frombuffer(raw[4 + 24 + n * 56:], dtype=np.float64)
    assert np.all(f2[:3] == ol.DEFAULT["cell_length"])
    rec = f2[3:].reshape(n, 7)
    x, y, z = o.positions()
    vx, vy, vz = o.velocities()
    assert np.array_equal(rec[:, 0], x) and np.array_equal(rec[:, 5], vz)
    assert np.array_equal(rec[:, 6], np.sqrt(vx * vx + vy * vy + vz * vz))


def test_brute_force_counts_agree_with_cell_list_counts():
    cfg = ol.config(4096)
    o = ol.Oracle(cfg)
    o.run(1, 120, DT)
    o2 = ol.Oracle(cfg, state=o.state())                 # fresh list on the melted state
    x, y, z = o2.positions()
    want = ol.oracle_pair_counts(cfg["cell_length"], x, y, z, cfg["rcut"] + cfg["shell"])
    assert np.array_equal(o2.neighbor_counts(), want)


def test_berendsen_pulls_temperature_and_nve_conserves_energy():
    cfg = ol.config(4096)
    o = ol.Oracle(cfg)
    o.run(1, 1500, DT)
    n = cfg["nr_particles"]
    t_kin = 2.0 * o.energies()[0] / (3.0 * (n - 1))
    assert abs(t_kin - cfg["kT"]) < 0.05
    o.set_tau(1e300)                                     # lambda rounds to exactly 1.0
    e = []
    for s in range(1501, 1901):
        o.integrate(s, DT)
        e.append(o.energies()[2])
    e = np.array(e)
    assert (e.max() - e.min()) / n < 2e-4
