"""Worker of the multi-device tests: one process per GPU (torchrun), gloo for the id broadcast.
Compares the z-slab path with the CPU oracle (and prints OK on rank 0)."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import oracle_lib as ol  # noqa: E402

DT = 0.001


def say(msg):
    if os.environ.get("LJGPU_WORKER_DEBUG"):
        print(f"[worker {os.environ.get('RANK')}] {msg}", file=sys.stderr, flush=True)


def main():
    n_atoms = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    long_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 150
    melt_steps = int(sys.argv[4]) if len(sys.argv) > 4 else 200
    min_rebuilds = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    lj = importlib.import_module("lennard-jones-live-simulator_b200")
    cfg = ol.config(n_atoms)
    # start from a melted state so atoms cross slab faces within a few hundred steps
    o = ol.Oracle(cfg)
    o.run(1, melt_steps, DT)
    state = o.state()
    o = ol.Oracle(cfg, state=state)

    uid = [lj.slab_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    params = lj.LennardJonesParameters(**cfg)
    g = lj.LennardJonesSimulation(params, state, force_kernel=lj.KERNEL_VERLET, device=local, slab=(rank, world, uid[0]))
    info = g.slab_info()
    say(f"created {info}")
    owned = torch.tensor([info["atoms_owned"]], dtype=torch.int64)
    dist.all_reduce(owned)
    assert int(owned.item()) == n_atoms, (int(owned.item()), n_atoms)

    def check(tag, tol=1e-12):
        ek, ep, et = g.get_energies()
        oek, oep, oet = o.energies()
        exact = o.epot_long_double()
        own = abs(oep - exact)
        assert abs(ep - oep) < tol * abs(oep) + 2 * own, (tag, ep, oep)
        if oek != 0.0:
            assert abs(ek - oek) < tol * abs(oek), (tag, ek, oek)
        err = ol.force_rel_error(g.get_forces_soa(), o.forces(), o.force_abs_sums()).max()
        assert err < tol, (tag, err)

    check("step 0")
    # the list every device walks, atom by atom, against the reference's list of the same positions, and the bit-exact
    # neighbour-count probe on the slab path (collective: every device assembles the whole system)
    assert np.array_equal(g.get_list_counts(), o.neighbor_counts())
    assert np.array_equal(g.get_neighbor_counts(cfg["rcut"] + cfg["shell"], True), o.neighbor_counts())
    assert np.array_equal(g.get_neighbor_counts(cfg["rcut"], False), o.incutoff_counts())
    for s in range(1, steps + 1):
        o.integrate(s, DT)
        g.integrate(s, DT)
        check(f"step {s}")
        say(f"step {s} ok")
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
    counts_mid = g.get_neighbor_counts(cfg["rcut"] + cfg["shell"], True)     # collective; atoms have moved since the binning
    if rank == 0 and n_atoms <= 65536:      # (the brute-force reference count is O(N^2))
        assert np.array_equal(counts_mid, ol.oracle_pair_counts(cfg["cell_length"], x, y, z, cfg["rcut"] + cfg["shell"]))
    dist.barrier()              # nobody steps on while rank 0 counts (a device waits LJGPU_PEER_TIMEOUT_S, default 30 s, for its peers, then reports them gone)

    # publication, the reference's per-step call pattern (integrate + publish_state): every device delivers its slice of
    # the original-order frame into ONE frame shared by all ranks; the frame must equal the collective getter bit for bit
    frames = lj.SharedFrames(n_atoms, world, rank, dist)
    ptrs = frames.pointers()
    pub_first = steps + 1
    for s in range(pub_first, pub_first + 6):
        o.integrate(s, DT)
        g.integrate_publish(s, DT, *ptrs[s & 1])
        if s > pub_first:                      # the delivery of step s - 1 is complete now, on this rank ...
            dist.barrier()                     # ... and after the barrier on every rank
            want = prev_positions
            got = frames.xyz((s - 1) & 1)
            for a, b in zip(got, want):
                assert np.array_equal(a, b), ("published frame", s - 1)
            dist.barrier()                     # nobody overwrites the frame while somebody compares
        prev_positions = g.get_positions_soa()
    g.publish_wait()
    dist.barrier()
    for a, b in zip(frames.xyz((pub_first + 5) & 1), prev_positions):
        assert np.array_equal(a, b), "last published frame"
    check(f"step {pub_first + 5} (after publishing)")
    dist.barrier()
    frames.close()
    steps = pub_first + 5

    # across re-binnings and migrations
    o.run(steps + 1, long_steps, DT)
    say("long run")
    dist.barrier()              # the oracle's run takes a different time on every rank (tens of seconds at N = 262 144)
    g.integrate_n(steps + 1, long_steps, DT)
    say("long run done")
    ek, ep, _ = g.get_energies()
    oek, oep, _ = o.energies()
    assert abs(ek - oek) < 1e-8 * abs(oek) and abs(ep - oep) < 1e-8 * abs(oep), (ek, oek, ep, oep)
    t = g.get_timers()
    assert t["rebuilds"] >= min_rebuilds
    owned = torch.tensor([g.slab_info()["atoms_owned"]], dtype=torch.int64)
    dist.all_reduce(owned)
    assert int(owned.item()) == n_atoms
    v = g.get_velocities_copy()
    assert np.abs(v.sum(axis=0)).max() < 1e-8
    dist.barrier()
    info = g.slab_info()
    g.close()

    # the pruned inner list (peer-memory step loop) must not change a single bit of the trajectory
    def variant(skin):
        if skin is None:
            os.environ.pop("LJGPU_PRUNE_SKIN", None)
        else:
            os.environ["LJGPU_PRUNE_SKIN"] = skin
        vid = [lj.slab_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(vid, src=0)
        s = lj.LennardJonesSimulation(params, state, force_kernel=lj.KERNEL_VERLET, device=local,
                                      slab=(rank, world, vid[0]))
        s.integrate_n(1, long_steps, DT)
        out = (s.get_energies(), s.get_positions_soa(), s.get_velocities_soa(), s.get_layout()["inner_list_steps"])
        dist.barrier()
        s.close()
        return out

    keep = os.environ.get("LJGPU_PRUNE_SKIN")
    e1, x1, v1, used1 = variant(keep)
    e0, x0, v0, used0 = variant("0")
    if keep is None:
        os.environ.pop("LJGPU_PRUNE_SKIN", None)
    else:
        os.environ["LJGPU_PRUNE_SKIN"] = keep
    assert used0 == 0
    assert e1 == e0, (e1, e0)
    for a, b in zip(x1 + v1, x0 + v0):
        assert np.array_equal(a, b)
    if not os.environ.get("LJGPU_SLAB_NCCL") and keep is None:
        assert used1 > 0, "inner list never used on the peer-memory path"
    dist.barrier()
    if rank == 0:
        print(f"SLAB OK world={world} N={n_atoms} rebuilds={t['rebuilds']} inner_steps={used1} info={info}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
