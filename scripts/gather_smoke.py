"""dev tool (torchrun, 2+ GPUs): smallest check of process_tile's gather transports against each other."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from nimrud_b200 import synth, distributed as nd
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
EDGES = (0.1, 0.2, 0.4, 0.8, 1.6); RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
cloud = synth.urban_scene(200_000, seed=31, device="cpu")
q = torch.quantile(cloud[:, 0].double(), torch.linspace(0, 1, world + 1, dtype=torch.float64))
mine = cloud[(cloud[:, 0] >= q[rank]) & ((cloud[:, 0] < q[rank + 1]) if rank + 1 < world else (cloud[:, 0] <= q[rank + 1]))].to(dev).contiguous()
local_rows = nd.process_tile(mine, EDGES, RADII)
out = torch.empty_like(local_rows)
a = nd.process_tile(mine, EDGES, RADII, gather=True, out=out)
b = nd.process_tile(mine, EDGES, RADII, gather="nccl")
c = nd.process_tile(mine, EDGES, RADII, gather="peer")
ok = bool(torch.equal(a, b) and torch.equal(a, c) and torch.equal(out, local_rows) and a.shape[0] == cloud.shape[0])
t = torch.tensor([int(ok)], device=dev); dist.all_reduce(t)
if rank == 0:
    print("gather smoke: world %d, %d rows, all transports identical and out= filled: %s" % (world, a.shape[0], bool(t.item() == world)), flush=True)
nd.release_mailboxes()
dist.destroy_process_group()
