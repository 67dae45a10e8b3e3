"""dev tool: kNN timing with queries in arbitrary order vs in the cell order of the feature kernels."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import _lib, multiscale, synth
from nimrud_b200._util import ptr, stream_ptr
n = 10_000_000
cloud = synth.urban_scene(n, seed=21, device="cuda")
index = multiscale.LatticeIndex(cloud, 0.1, indexed=True)
lib = _lib.lib()
lo = cloud.min(0).values.double().cpu().numpy(); hi = cloud.max(0).values.double().cpu().numpy()
box = np.concatenate([lo, hi]); f64p = ctypes.POINTER(ctypes.c_double)
origin = np.zeros(3)
_lib.check(lib.nbr_brick_origin(box.ctypes.data_as(f64p), None, 0.1, origin.ctypes.data_as(f64p)))
perm = torch.empty(n, dtype=torch.int32, device="cuda"); ordered = torch.empty_like(cloud)
_lib.check(lib.nbr_order_cloud(ptr(cloud), _lib.F32, n, box.ctypes.data_as(f64p), origin.ctypes.data_as(f64p), 0.1, ptr(perm), ptr(ordered), stream_ptr(cloud.device)))
for name, q in (("arbitrary order", cloud), ("cell order", ordered)):
    for k, ks in ((10, (10,)), (50, (10, 20, 50))):
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            res = index.knn(q, k, ks=ks, out_dtype=np.float32)
            e1.record(); torch.cuda.synchronize(); dt = e0.elapsed_time(e1) * 1e-3
        print("%s k=%d: %.1f ms -> %.2f M queries/s" % (name, k, dt * 1e3, n / dt / 1e6))
        del res
