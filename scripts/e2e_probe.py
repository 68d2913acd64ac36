"""Developer probe (under torchrun): where the per-call time of the drop-in call pattern goes on the slab path."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
lj = importlib.import_module("lennard-jones-live-simulator_b200")
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
L = float(np.cbrt(n / 0.8))
p = lj.LennardJonesParameters(cell_length=L, nr_particles=n, kT=1.0, mass=1.0, sigma=1.0, rcut=2.5, epsilon=1.0, tau=0.5, shell=0.4, stepsize=0.001)
lj.host_rng_reset(); st = lj.host_initialize(p)
if world > 1:
    uid = [lj.slab_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0, device=torch.device("cpu"))
    sim = lj.LennardJonesSimulation(p, st, force_kernel=lj.KERNEL_VERLET, device=lr, slab=(rank, world, uid[0]))
else:
    sim = lj.LennardJonesSimulation(p, st, force_kernel=lj.KERNEL_VERLET, device=lr)
step = 1
sim.integrate_n(step, 500, 0.001); step += 500
def bar():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
K = 200
def timed(name, fn):
    global step
    bar(); t0 = time.perf_counter()
    for _ in range(K):
        fn(); step += 1
    sim.publish_wait(); bar()
    if rank == 0: print(f"{name:50s} {(time.perf_counter() - t0) / K * 1e6:8.1f} us/call", flush=True)
shm = lj.SharedFrames(n, world, rank, dist if world > 1 else None)
ptrs = shm.pointers()
bar(); t0 = time.perf_counter(); sim.integrate_n(step, K, 0.001); step += K; bar()
if rank == 0: print(f"{'integrate_n(K) / K':50s} {(time.perf_counter() - t0) / K * 1e6:8.1f} us/step", flush=True)
timed("integrate(1)", lambda: sim.integrate(step, 0.001))
timed("integrate(1) + get_energies", lambda: (sim.integrate(step, 0.001), sim.get_energies()))
timed("integrate_publish", lambda: sim.integrate_publish(step, 0.001, *ptrs[step & 1]))
timed("integrate_publish + get_energies", lambda: (sim.integrate_publish(step, 0.001, *ptrs[step & 1]), sim.get_energies()))
shm.close(); sim.close()
