"""BASELINE config 4: 100M-point synthetic aerial-LiDAR tile, 5 scales (edges 0.5..8 m, r = 3e), spatially sharded
with halo exchange (torchrun, one rank per GPU; the 100M points are split into world_size tiles side by side).
prints points*scales/s, and on rank 0 checks a sample of rows against the CPU oracle."""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from nimrud_b200 import distributed as nd, synth
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
TOTAL = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
EDGES = (0.5, 1.0, 2.0, 4.0, 8.0); RADII = (1.5, 3.0, 6.0, 12.0, 24.0)
n = TOTAL // world
extent = math.sqrt(n / 8.0)
cols = 2 if world >= 2 else 1
cloud = synth.aerial_tile(n, seed=22 + rank, device=dev, origin=((rank % cols) * extent, (rank // cols) * extent))
out = torch.empty((n, 20), dtype=torch.float32, device=dev)
for it in range(4):
    dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
    nd.process_tile(cloud, EDGES, RADII, out=out, gather=False)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0 and it > 0:
        print("step %d: %.2f ms, %.2f G point*scales/s (%d GPUs, %d points)" % (it, dt * 1e3, n * world * 5 / dt / 1e9, world, n * world), flush=True)
all_boxes, _ = nd.gather_boxes(cloud)
if rank == 0:
    # rows of this tile far from its border only need nearby points: check a sample against the CPU oracle, with the
    # oracle's voxel grids anchored on the GLOBAL box as the tiles are (utils/geometry.py:37-62 on the whole cloud)
    from oracle import c_oracle
    c_oracle.build()
    g_lo = all_boxes[:, :3].min(0).values.numpy(); g_hi = all_boxes[:, 3:].max(0).values.numpy()
    lo = cloud.min(0).values; hi = cloud.max(0).values
    centre = ((lo + hi) / 2).cpu().numpy()
    xy = cloud[:, :2].cpu().numpy()
    near = np.abs(xy - centre[:2]).max(1) < 150.0
    sub = cloud.cpu().numpy()[near].astype(np.float64)
    core = np.abs(sub[:, :2] - centre[:2]).max(1) < 60.0
    q = sub[core][:3000]
    rows = np.nonzero(near)[0][core][:3000]
    got = out[torch.from_numpy(rows).to(dev)].cpu().numpy().astype(np.float64)
    for s_, (e, r) in enumerate(zip(EDGES, RADII)):
        minc = g_lo - e / 2
        widths = np.ceil(np.log2(((g_hi + e / 2) - minc) / e)).astype(np.int64)
        ukeys, _ = c_oracle.unique_voxels(sub, minc, e, widths)
        ref = c_oracle.radius_features(q, ukeys, minc, e, widths, r)
        assert np.array_equal(got[:, 4 * s_], ref[:, 0]), "populations differ at scale %d" % s_
        assert np.abs(got[:, 4 * s_ + 1] - ref[:, 1]).max() <= 1e-4 * r
        for j in (2, 3):
            d = np.abs(got[:, 4 * s_ + j] - ref[:, j])
            assert (d <= 1e-4 * np.abs(ref[:, j]) + 1e-9).all()
    print("oracle check ok on %d rows; mean populations %s" % (len(q), [round(float(got[:, 4 * s].mean()), 1) for s in range(5)]), flush=True)
dist.destroy_process_group()
