"""dev tool (torchrun, one rank per GPU): host-buffer tile step with the routes of a pinned float64 result
(NBR_HOST_WIDEN = auto / host / device, given as arguments)."""
import math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from nimrud_b200 import synth, distributed as nd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 10_000_000
EDGES = (0.1, 0.2, 0.4, 0.8, 1.6); RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
extent = math.sqrt(n / 40.0)
cloud = synth.urban_scene(n, seed=20 + rank, device=dev, origin=((rank % 2) * extent, (rank // 2) * extent))
host_in = cloud.cpu().pin_memory()
host_out = torch.empty((n, 20), dtype=torch.float64).pin_memory()
for share in sys.argv[1:]:
    os.environ["NBR_HOST_WIDEN"] = share
    nd.process_tile_host(host_in, EDGES, RADII, out=host_out, out_dtype=np.float64, device=dev)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        nd.process_tile_host(host_in, EDGES, RADII, out=host_out, out_dtype=np.float64, device=dev)
    torch.cuda.synchronize()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("world %d route %s: %.1f ms per step, %.3f G point*scales/s" % (world, share, t.item() / 3 * 1e3, world * n * 5 * 3 / t.item() / 1e9), flush=True)
nd.release_mailboxes()
dist.destroy_process_group()
