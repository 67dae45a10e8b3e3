"""dev tool: the radix sort + unique + indexed lattice build of the kNN / neighbor-index path, CUDA-event timing.
prints the achieved bandwidth of the sort against its algorithmic bytes (one read + one write of the keys per pass)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import _lib, multiscale, synth
from nimrud_b200.geometry import VoxelFilter
lib = _lib.lib()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cloud = synth.urban_scene(n, seed=21, device="cuda")
vf = VoxelFilter(cloud, 0.1)
bits = int(sum(vf.widths))
addr = vf.coordinate_to_address(cloud)
keys0 = addr.to(torch.int64).contiguous()
tmp = torch.empty_like(keys0)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
keys = keys0.clone()
def sort_once():
    keys.copy_(keys0)
    _lib.check(lib.nbr_sort_u64(ctypes.c_void_p(keys.data_ptr()), ctypes.c_void_p(tmp.data_ptr()), n, 0, bits,
                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
t_copy = timed(lambda: keys.copy_(keys0))
t_sort = timed(sort_once) - t_copy
passes = (bits + 7) // 8
print("radix sort of %d 64-bit keys, %d key bits (%d passes): %.3f ms -> %.1f GB/s of algorithmic traffic (16 B per key and pass)"
      % (n, bits, passes, t_sort, n * 16 * passes / t_sort / 1e6))
assert torch.equal(keys, torch.sort(keys0).values)
def build():
    ix = multiscale.LatticeIndex(cloud, 0.1, indexed=True, bbox=box)
    ix.close()
from nimrud_b200.geometry import cloud_bbox
box = cloud_bbox(cloud, _lib.F32)
print("indexed lattice build (bit bricks + addresses + sort + unique + row ranks), 10M points: %.3f ms" % timed(build, 3))
u = vf.unique_voxels(cloud)
print("unique voxels:", u.shape[0])
