"""Development aid: device-to-host bandwidth into cudaMallocHost memory and into a registered shared-memory mapping
(pipeline.SharedHostStream), one GPU; with torchrun, every rank copies its share into the same shared buffer at once."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from terminalraytracer_b200 import pipeline, renderer as R
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
n = 829444329
rd = R.Renderer(torch.cuda.current_device())
d = torch.empty(n, dtype=torch.uint8, device="cuda")
shared = pipeline.SharedHostStream(rd, n, rank, world)
share = n // world
lo = rank * share


def bw(dst_ptr, label):
    L = rd.L
    for _ in range(2):
        L.trt_push_to_peer(dst_ptr + lo, d.data_ptr() + lo, share); L.trt_peer_copies_wait()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        L.trt_push_to_peer(dst_ptr + lo, d.data_ptr() + lo, share)
    L.trt_peer_copies_wait()
    if world > 1:
        dist.barrier()
    dt = (time.perf_counter() - t0) / 5
    if rank == 0:
        print("%s: %d ranks x %.1f MB in %.2f ms = %.1f GB/s aggregate" % (label, world, share / 1e6, dt * 1e3, share * world / dt / 1e9), flush=True)


bw(shared.ptr, "registered shared memory")
if world == 1:
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    bw(h.data_ptr(), "cudaMallocHost")
shared.close()
rd.close()
if world > 1:
    dist.destroy_process_group()
