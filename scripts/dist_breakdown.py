"""dev tool: wall-clock breakdown of nimrud_b200.distributed.process_tile on every rank (torchrun)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, numpy as np, torch, torch.distributed as dist
from nimrud_b200 import distributed as nd, synth, multiscale
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
EDGES = (0.1, 0.2, 0.4, 0.8, 1.6); RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
n = 10_000_000; extent = math.sqrt(n / 40.0); cols = 2
cloud = synth.urban_scene(n, seed=20 + rank, device=dev, origin=((rank % cols) * extent, (rank // cols) * extent))
out = torch.empty((n, 20), dtype=torch.float32, device=dev)
def sync(): torch.cuda.synchronize()
for it in range(6):
    dist.barrier(); sync(); t0 = time.perf_counter()
    halo, (g_lo, g_hi), _ = nd.exchange_halo(cloud, EDGES, RADII); sync(); t1 = time.perf_counter()
    search = torch.cat([cloud, halo], 0); bbox = (g_lo.cpu().numpy(), g_hi.cpu().numpy()); sync(); t2 = time.perf_counter()
    multiscale.process_single_core(search[:n], search, EDGES, RADII, out_dtype=np.float32, global_bbox=bbox, out=out); sync(); t3 = time.perf_counter()
    if it >= 3:
        print("rank %d: exchange %.2f ms (halo %d pts), cat+bbox %.2f ms, features call %.2f ms" % (rank, (t1 - t0) * 1e3, halo.shape[0], (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
dist.destroy_process_group()
