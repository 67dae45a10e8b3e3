"""dev tool: time kNN (config 3 shape) on the GPU with CUDA events."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import multiscale, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cloud = synth.urban_scene(n, seed=21, device="cuda")
t0 = time.perf_counter(); index = multiscale.LatticeIndex(cloud, 0.1, indexed=True); nv = index.n_voxels; torch.cuda.synchronize()
print("indexed lattice build %.1f ms, %d voxels" % ((time.perf_counter() - t0) * 1e3, nv))
for k, ks in ((10, None), (10, (10,)), (50, (10, 20, 50))):
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        res = index.knn(cloud, k, ks=ks, out_dtype=np.float32)
        e1.record(); torch.cuda.synchronize(); dt = e0.elapsed_time(e1) * 1e-3
    print("k=%d ks=%s: %.1f ms for %d queries -> %.2f M queries/s" % (k, ks, dt * 1e3, n, n / dt / 1e6))
    del res
