"""dev tool (torchrun, one rank per GPU): the tile step with the feature all-gather -- NCCL, peer stores from the feature
kernel (fused), peer copies of the finished share (copy)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from nimrud_b200 import synth, distributed as nd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
EDGES = (0.1, 0.2, 0.4, 0.8, 1.6); RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
extent = math.sqrt(n / 40.0)
cloud = synth.urban_scene(n, seed=20 + rank, device=dev, origin=((rank % 2) * extent, (rank // 2) * extent))
def timed(label, fn, steps=5):
    fn(); fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("world %d n %d %-28s %.2f ms per step" % (world, n, label, t.item()), flush=True)
out = torch.empty((n, 20), dtype=torch.float32, device=dev)
timed("no gather", lambda: nd.process_tile(cloud, EDGES, RADII, out=out))
timed("nccl all-gather", lambda: nd.process_tile(cloud, EDGES, RADII, out=out, gather="nccl"))
res = torch.empty((world * n, 20), dtype=torch.float32, device=dev)
for mode in sys.argv[2:] or ["fused", "copy"]:
    os.environ["NBR_GATHER"] = mode
    timed("peer gather (%s)" % mode, lambda: nd.process_tile(cloud, EDGES, RADII, gather="peer", out_all=res))
nd.release_mailboxes()
dist.destroy_process_group()
