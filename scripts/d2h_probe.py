"""Developer probe (under torchrun): aggregate device-to-host bandwidth of `world` GPUs delivering slices of one frame at the same
time, into (a) each process's own page-locked buffer, (b) one POSIX shared-memory frame registered with CUDA by every process."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
from multiprocessing import shared_memory
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", lr))
total = 3 * 1048576 * 8
mine = total // world
src = torch.empty(mine, dtype=torch.uint8, device="cuda")
own = torch.empty(mine, dtype=torch.uint8).pin_memory()
names = [None]
if rank == 0:
    seg = shared_memory.SharedMemory(create=True, size=total); names[0] = seg.name
dist.broadcast_object_list(names, src=0, device=torch.device("cpu"))
if rank != 0: seg = shared_memory.SharedMemory(name=names[0])
arr = np.ndarray((total,), dtype=np.uint8, buffer=seg.buf)
if rank == 0: arr[:] = 0
dist.barrier()
rt = torch.cuda.cudart()
rc = rt.cudaHostRegister(arr.ctypes.data, total, 0)
assert int(rc) == 0, rc
shm_t = torch.from_numpy(arr)[rank * mine:(rank + 1) * mine]
def run(name, dst):
    for _ in range(5): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(50): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dist.barrier()
    dt = (time.perf_counter() - t0) / 50
    if rank == 0: print(f"{name:40s} world={world}: {dt*1e6:8.1f} us per frame of {total/1e6:.1f} MB = {total/dt/1e9:6.1f} GB/s aggregate", flush=True)
run("own page-locked buffers", own)
run("registered POSIX shared memory", shm_t)
seg.close()
if rank == 0: seg.unlink()
