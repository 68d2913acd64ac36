"""Developer probe (under torchrun): the slab path alone from a melted state - with the pruned inner list and without it
(LJGPU_PRUNE_SKIN=0), in one process like tests/slab_worker.py - and compare."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
lj = importlib.import_module("lennard-jones-live-simulator_b200")
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]); steps = int(sys.argv[2]); melt = int(sys.argv[3]) if len(sys.argv) > 3 else 0
order = sys.argv[4].split(",") if len(sys.argv) > 4 else ["default", "0"]
L = float(np.cbrt(n / 0.8))
p = lj.LennardJonesParameters(cell_length=L, nr_particles=n, kT=1.0, mass=1.0, sigma=1.0, rcut=2.5, epsilon=1.0, tau=0.5, shell=0.4, stepsize=0.001)
lj.host_rng_reset(); st = lj.host_initialize(p)
if melt:
    m = lj.LennardJonesSimulation(p, st, force_kernel=lj.KERNEL_VERLET, device=lr)
    m.integrate_n(1, melt, 0.001)
    x, y, z = m.get_positions_soa(); vx, vy, vz = m.get_velocities_soa()
    st = (x, y, z, vx, vy, vz)
    m.close()
res = {}
for skin in order:
    if skin == "default": os.environ.pop("LJGPU_PRUNE_SKIN", None)
    else: os.environ["LJGPU_PRUNE_SKIN"] = skin
    uid = [lj.slab_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, device=torch.device("cpu"))
    g = lj.LennardJonesSimulation(p, st, force_kernel=lj.KERNEL_VERLET, device=lr, slab=(rank, world, uid[0]))
    g.integrate_n(1, steps, 0.001)
    res[skin] = g.get_energies()
    if rank == 0: print(f"  skin {skin}: {res[skin]} rebuilds {g.get_timers()['rebuilds']}", flush=True)
    dist.barrier()
    g.close()
if rank == 0: print("REPRO", "OK" if len(set(res.values())) == 1 else "MISMATCH", flush=True)
