"""dev tool: wall-clock breakdown of the CUDA tile path of nimrud_b200.distributed on every rank (torchrun)."""
import os, sys, time, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import math, numpy as np, torch, torch.distributed as dist
from nimrud_b200 import distributed as nd, synth, multiscale, _lib
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
    g = nd.gather_boxes(cloud); sync(); t1 = time.perf_counter()
    halo, _, _ = nd.exchange_halo(cloud, EDGES, RADII, gathered=g); sync(); t2 = time.perf_counter()
    dist.barrier(); sync(); t3 = time.perf_counter()
    nd.process_tile(cloud, EDGES, RADII, out=out); sync(); t4 = time.perf_counter()
    if it >= 3:
        print("rank %d: gather_boxes %.2f ms, exchange %.2f ms (halo %d pts), whole process_tile %.2f ms" % (
            rank, (t1 - t0) * 1e3, (t2 - t1) * 1e3, halo.shape[0], (t4 - t3) * 1e3), flush=True)
dist.destroy_process_group()
