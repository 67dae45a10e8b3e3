"""dev tool (torchrun, one rank per GPU): device->host and host->device bandwidth of 800 MB pinned copies issued by all
ranks at once -- the ceiling of the host-buffer path at N GPUs of one box."""
import os, time, torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
dev = torch.empty(200_000_000, dtype=torch.float32, device="cuda")
host = torch.empty(200_000_000, dtype=torch.float32).pin_memory()
for name, fn in (("d2h", lambda: host.copy_(dev, non_blocking=True)), ("h2d", lambda: dev.copy_(host, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("%s: %d ranks x 0.8 GB in %.2f ms (slowest rank) -> %.1f GB/s per rank, %.1f GB/s aggregate" % (
            name, world, t.item() * 1e3, 0.8 / t.item(), 0.8 * world / t.item()), flush=True)
if world > 1: dist.destroy_process_group()
