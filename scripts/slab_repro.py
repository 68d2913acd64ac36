"""Developer probe (under torchrun): the slab path alone - create, run `steps` steps in one call, print the energies."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
lj = importlib.import_module("lennard-jones-live-simulator_b200")
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(lr)
dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]); steps = int(sys.argv[2])
L = float(np.cbrt(n / 0.8))
p = lj.LennardJonesParameters(cell_length=L, nr_particles=n, kT=1.0, mass=1.0, sigma=1.0, rcut=2.5, epsilon=1.0, tau=0.5, shell=0.4, stepsize=0.001)
lj.host_rng_reset(); st = lj.host_initialize(p)
uid = [lj.slab_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0, device=torch.device("cpu"))
g = lj.LennardJonesSimulation(p, st, force_kernel=lj.KERNEL_VERLET, device=lr, slab=(rank, world, uid[0]))
t0 = time.time()
done = 0
while done < steps:
    k = min(200, steps - done)
    g.integrate_n(done + 1, k, 0.001); done += k
    e = g.get_energies()
    if rank == 0: print(f"  steps {done}: ekin {e[0]:.6f} epot {e[1]:.6f}  ({time.time() - t0:.1f} s)", flush=True)
dist.barrier()
if rank == 0: print("REPRO OK", g.slab_info(), flush=True)
g.close()
