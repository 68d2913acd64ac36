"""CPU tests of the slab path's host-side logic with two gloo ranks: the layer partition of
ljgpu_slab_layers is a disjoint cover, ring neighbours agree on the layers they exchange, the owner
rule used at start-up assigns every atom to exactly one rank, and between two re-binnings of the
reference atoms only ever move to an adjacent layer (the assumption behind one-neighbour migration)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nc_list, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as ol
    lj = importlib.import_module("lennard-jones-live-simulator_b200")
    try:
        for nc in nc_list:
            lo, hi = lj.slab_layers(nc, world, rank)
            mine = [None] * world
            dist.all_gather_object(mine, (lo, hi))
            # disjoint, ordered cover of [0, nc)
            assert mine[0][0] == 0 and mine[-1][1] == nc
            for a, b in zip(mine, mine[1:]):
                assert a[1] == b[0]
            sizes = [b - a for a, b in mine]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
            # ring neighbours: my upper halo layer is the up-neighbour's first layer, and vice versa
            up, down = (rank + 1) % world, (rank + world - 1) % world
            assert hi % nc == mine[up][0] % nc
            assert (lo + nc - 1) % nc == (mine[down][1] - 1) % nc

        # owner rule on a real configuration: every atom has exactly one owner
        cfg = ol.config(4096)
        o = ol.Oracle(cfg)
        o.run(1, 150, 0.001)
        cells, dims = o.cell_ids()
        nc = dims[0]
        layer = cells // (nc * nc)
        lo, hi = lj.slab_layers(nc, world, rank)
        owned = (layer >= lo) & (layer < hi)
        counts = [None] * world
        dist.all_gather_object(counts, int(owned.sum()))
        assert sum(counts) == cfg["nr_particles"]
        # between two list rebuilds of the reference no atom moves further than one layer
        before = layer.copy()
        r0 = o.rebuild_count()
        for s in range(151, 400):
            o.integrate(s, 0.001)
            if o.rebuild_count() != r0:
                now = o.cell_ids()[0] // (nc * nc)
                d = (now.astype(np.int64) - before.astype(np.int64)) % nc
                assert np.all((d == 0) | (d == 1) | (d == nc - 1))
                before, r0 = now, o.rebuild_count()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_slab_partition_and_owner_rule_two_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, [3, 5, 14, 37, 93], q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_slab_layers_all_world_sizes():
    lj = importlib.import_module("lennard-jones-live-simulator_b200")
    for nc in (3, 5, 14, 37, 93):
        for world in (1, 2, 3, 4, 8):
            if world > nc:
                continue
            spans = [lj.slab_layers(nc, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nc
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(lj.LJGpuError):
        lj.slab_layers(5, 2, 2)
