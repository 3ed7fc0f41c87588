"""D2H bandwidth microbenchmark at N ranks: every rank copies its band set (12.4 MB at 8 ranks of a 4K film) from
device memory to (a) its own cudaHostAlloc'ed buffer, (b) a POSIX shared-memory frame registered with
cudaHostRegister (what vrt_render_bands_async targets).  Run under torchrun; prints GB/s per rank and aggregate."""
import os, sys, time
import numpy as np, torch, torch.distributed as td
sys.path.insert(0, '.')
from voxelraytrace20190722_b200 import capi, dist as vdist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
td.init_process_group("nccl", device_id=torch.device("cuda", local))
capi.load()
ny, nx = 2160, 3840
nbytes = ny * nx * 12 // world
src = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
own = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
shf = vdist.SharedHostFrame(ny, nx, nbuf=2)
shared = torch.from_numpy(np.frombuffer(shf._mm, dtype=np.uint8))[rank * nbytes:(rank + 1) * nbytes]
def run(dst, reps=200):
    torch.cuda.synchronize(); td.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); td.barrier()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9
for name, dst in (("cudaHostAlloc (own buffer)", own), ("shm + cudaHostRegister (shared frame)", shared)):
    run(dst, 20)
    g = torch.tensor([run(dst)], device="cuda")
    td.all_reduce(g)
    if rank == 0:
        print(f"{name}: {float(g) / world:.1f} GB/s per rank, {float(g):.1f} GB/s aggregate over {world} ranks", flush=True)
td.destroy_process_group()
