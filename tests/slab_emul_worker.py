"""Worker of tests/test_emulated_library.py::test_slab_path_on_the_emulated_library: the z-slab path with the "devices"
as THREADS of this process, against the CPU emulation of the product library (tests/cuda_emul) and its stand-in NCCL.
Every rank makes the same collective calls and compares what it gets with its own oracle.  Mirrors
tests/slab_worker.py (the multi-GPU worker).  Test infrastructure only."""
import importlib
import os
import sys
import threading
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as ol  # noqa: E402

DT = 0.001


def main():
    world = int(sys.argv[1])
    n_atoms = int(sys.argv[2])
    steps, long_steps, melt_steps, min_rebuilds = (int(a) for a in sys.argv[3:7])
    lj = importlib.import_module("lennard-jones-live-simulator_b200")
    cfg = ol.config(n_atoms)
    o0 = ol.Oracle(cfg)
    o0.run(1, melt_steps, DT)
    state = o0.state()
    params = lj.LennardJonesParameters(**cfg)
    bar = threading.Barrier(world)
    shared = {"uid": lj.slab_unique_id(), "owned": [0] * world, "out": [None] * world}
    errors = []

    def body(rank):
        o = ol.Oracle(cfg, state=state)
        g = lj.LennardJonesSimulation(params, state, force_kernel=lj.KERNEL_VERLET, device=rank, slab=(rank, world, shared["uid"]))

        def owned_total():
            shared["owned"][rank] = g.slab_info()["atoms_owned"]
            bar.wait()
            tot = sum(shared["owned"])
            bar.wait()
            return tot

        assert owned_total() == n_atoms

        def check(tag, tol=1e-12):
            ek, ep, et = g.get_energies()
            oek, oep, oet = o.energies()
            own = abs(oep - o.epot_long_double())
            assert abs(ep - oep) < tol * abs(oep) + 2 * own, (tag, ep, oep)
            if oek != 0.0:
                assert abs(ek - oek) < tol * abs(oek), (tag, ek, oek)
            err = ol.force_rel_error(g.get_forces_soa(), o.forces(), o.force_abs_sums()).max()
            assert err < tol, (tag, err)

        if os.environ.get("LJGPU_TEST_FAIL_RANK"):
            # fault injection: one device "runs out of halo space" at its second re-binning - EVERY device must come back
            # with the same LJGPU_ESTATE error instead of waiting forever for the one that gave up
            try:
                g.integrate_n(1, 400, DT)
                shared["out"][rank] = "no error"
            except lj.LJGpuError as e:
                shared["out"][rank] = str(e)
            bar.wait()
            g.close()
            return
        check("step 0")
        assert np.array_equal(g.get_list_counts(), o.neighbor_counts())          # the list every "device" walks
        assert np.array_equal(g.get_neighbor_counts(cfg["rcut"] + cfg["shell"], True), o.neighbor_counts())
        for s in range(1, steps + 1):
            o.integrate(s, DT)
            g.integrate(s, DT)
            check(f"step {s}")
        L = cfg["cell_length"]
        for a, b in zip(g.get_positions_soa(), o.positions()):
            d = np.abs(a - b); d = np.minimum(d, L - d)
            assert d.max() < 1e-12
        for a, b in zip(g.get_velocities_soa(), o.velocities()):
            assert np.abs(a - b).max() < 1e-12
        cells, dims = g.get_cell_ids()
        x, y, z = g.get_positions_soa()
        assert np.array_equal(cells, ol.oracle_cell_ids_of(cfg, x, y, z)[0])
        vx, vy, vz = g.get_velocities_soa()
        assert np.array_equal(g.get_speed_histogram(), ol.oracle_speed_histogram(vx, vy, vz))
        counts_mid = g.get_neighbor_counts(cfg["rcut"], True)                    # collective; mid-life of a list
        if rank == 0:
            assert np.array_equal(counts_mid, ol.oracle_pair_counts(cfg["cell_length"], x, y, z, cfg["rcut"]))
        bar.wait()              # nobody steps on while rank 0 counts: a device waits ~2 s for its peers, then reports them gone

        # publication: every "device" delivers its slice of the original-order frame into one frame shared by all ranks
        bar.wait()
        if rank == 0:
            shared["frames"] = [np.zeros(3 * n_atoms), np.zeros(3 * n_atoms)]
        bar.wait()
        fr = shared["frames"]
        ptrs = [[f.ctypes.data + k * n_atoms * 8 for k in range(3)] for f in fr]
        first_pub = steps + 1
        for s in range(first_pub, first_pub + 4):
            o.integrate(s, DT)
            g.integrate_publish(s, DT, *ptrs[s & 1])
        g.publish_wait()
        bar.wait()
        last = first_pub + 3
        want = g.get_positions_soa()
        f = fr[last & 1]
        for k in range(3):
            assert np.array_equal(f[k * n_atoms:(k + 1) * n_atoms], want[k]), "published frame differs from the getter"
        check(f"step {last} (after publishing)")
        bar.wait()
        steps_done = last

        o.run(steps_done + 1, long_steps, DT)                  # across re-binnings and migrations
        g.integrate_n(steps_done + 1, long_steps, DT)
        ek, ep, _ = g.get_energies()
        oek, oep, _ = o.energies()
        assert abs(ek - oek) < 1e-8 * abs(oek) and abs(ep - oep) < 1e-8 * abs(oep), (ek, oek, ep, oep)
        rebuilds = g.get_timers()["rebuilds"]
        assert rebuilds >= min_rebuilds, rebuilds
        assert owned_total() == n_atoms
        v = g.get_velocities_copy()
        assert np.abs(v.sum(axis=0)).max() < 1e-8
        info = g.slab_info()
        bar.wait()
        g.close()

        def variant(skin):                                      # the pruned inner list must not change a single bit
            bar.wait()
            if rank == 0:
                if skin is None:
                    os.environ.pop("LJGPU_PRUNE_SKIN", None)
                else:
                    os.environ["LJGPU_PRUNE_SKIN"] = skin
                shared["uid"] = lj.slab_unique_id()
            bar.wait()
            s = lj.LennardJonesSimulation(params, state, force_kernel=lj.KERNEL_VERLET, device=rank, slab=(rank, world, shared["uid"]))
            s.integrate_n(1, long_steps, DT)
            out = (s.get_energies(), s.get_positions_soa(), s.get_velocities_soa(), s.get_layout()["inner_list_steps"])
            bar.wait()
            s.close()
            return out

        e1, x1, v1, used1 = variant(None)
        e0, x0, v0, used0 = variant("0")
        assert used0 == 0 and e1 == e0, (used0, e1, e0)
        for a, b in zip(x1 + v1, x0 + v0):
            assert np.array_equal(a, b)
        if not os.environ.get("LJGPU_SLAB_NCCL"):
            assert used1 > 0, "inner list never used on the peer-memory path"
        shared["out"][rank] = (rebuilds, used1, info)

    def guarded(rank):
        try:
            body(rank)
        except BaseException:                                    # a failing rank must not leave the others in a barrier
            errors.append((rank, traceback.format_exc()))
            bar.abort()
            os._exit(1) if False else None

    threads = [threading.Thread(target=guarded, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=1200)
    if errors or any(t.is_alive() for t in threads):
        for rank, tb in errors:
            print(f"--- rank {rank} ---\n{tb}", file=sys.stderr)
        sys.stderr.flush()
        os._exit(1)                                             # ranks stuck in a collective cannot be joined
    if os.environ.get("LJGPU_TEST_FAIL_RANK"):
        msgs = shared["out"]
        assert all(isinstance(m, str) and "slab path: a halo layer does not fit" in m for m in msgs), msgs
        print(f"SLAB EMUL FAILED TOGETHER world={world}: {msgs[0]}")
        return
    rebuilds, used, info = shared["out"][0]
    print(f"SLAB EMUL OK world={world} N={n_atoms} rebuilds={rebuilds} inner_steps={used} info={info}")


if __name__ == "__main__":
    main()
