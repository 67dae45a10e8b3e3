"""dev tool: time kNN (config 3 shape) on the GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import multiscale, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cloud = synth.urban_scene(n, seed=21, device="cuda")
t0 = time.perf_counter(); index = multiscale.LatticeIndex(cloud, 0.1, indexed=True); nv = index.n_voxels; torch.cuda.synchronize()
print("indexed lattice build %.1f ms, %d voxels" % ((time.perf_counter() - t0) * 1e3, nv))
for k, ks in ((10, (10,)), (50, (10, 20, 50))):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        idx, d2, f = index.knn(cloud, k, ks=ks, out_dtype=np.float32)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("k=%d: %.1f ms for %d queries -> %.2f M queries/s" % (k, dt * 1e3, n, n / dt / 1e6))
