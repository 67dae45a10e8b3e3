"""dev tool: raw-point kNN timing (config 3 shape) with CUDA events, and the same kernel on the voxel centres."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import multiscale, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cloud = synth.urban_scene(n, seed=21, device="cuda")
q = synth.with_ties(cloud, 0.1, seed=21, fraction=0.01)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for edge in (0.0, 0.2, 0.3, 0.45, 0.6):
    for k, ks in ((10, (10,)), (50, (10, 20, 50))):
        ms = timed(lambda: multiscale.knn_points(q, cloud, k, ks=ks, out_dtype=np.float32, cell_edge=edge))
        print("raw points, cell edge %.2f, k=%d: %.1f ms -> %.1f M queries/s" % (edge, k, ms, q.shape[0] / ms / 1e3), flush=True)
index = multiscale.LatticeIndex(cloud, 0.1, indexed=True)
_, cen = index.addresses_and_centres()
ms = timed(lambda: multiscale.knn_points(q, cen, 50, ks=(10, 20, 50), out_dtype=np.float32))
print("voxel centres through the point kernel, k=50: %.1f ms -> %.1f M queries/s" % (ms, q.shape[0] / ms / 1e3))
ms = timed(lambda: index.knn(q, 50, ks=(10, 20, 50), out_dtype=np.float32))
print("voxel centres through the brick kernel, k=50: %.1f ms -> %.1f M queries/s" % (ms, q.shape[0] / ms / 1e3))
a = multiscale.knn_points(q[:200000], cen, 50)
b = index.knn(q[:200000], 50)
print("identical index sets and distances:", bool(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])))
