"""dev tool: equal-edge multi-radius variant of config 2 and config 1, CUDA-event timing."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import _lib, multiscale, synth
lib = _lib.lib(); lib.nbr_timing_enable(1); ph = (ctypes.c_double * 8)()
def run(name, cloud, edges, radii, reps=3):
    for _ in range(2):
        out = multiscale.process_single_core(cloud, cloud, edges, radii, out_dtype=np.float32)
    torch.cuda.synchronize(); lib.nbr_timing_read(ph)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = multiscale.process_single_core(cloud, cloud, edges, radii, out_dtype=np.float32)
    e1.record(); torch.cuda.synchronize(); lib.nbr_timing_read(ph)
    ms = e0.elapsed_time(e1) / reps
    n = cloud.shape[0]
    print("%s: %.3f ms/step, %.3f G pt*scales/s; phases bbox %.3f index %.3f order %.3f features %.3f; mean pops %s" % (
        name, ms, n * len(radii) / ms / 1e6, ph[0] / reps, ph[1] / reps, ph[2] / reps, ph[3] / reps,
        [round(float(out[:, 4 * s].mean()), 1) for s in range(len(radii))]))
cloud = synth.urban_scene(10_000_000, seed=20, device="cuda")
run("config2 (r = 3e)", cloud, [0.1, 0.2, 0.4, 0.8, 1.6], [0.3, 0.6, 1.2, 2.4, 4.8])
run("config2 equal-edge e=0.2, r=0.4..1.2", cloud, [0.2] * 5, [0.4, 0.6, 0.8, 1.0, 1.2])
run("config2 equal-edge e=0.2, r=0.4,0.6 only", cloud, [0.2] * 2, [0.4, 0.6])
rs = np.random.RandomState(10)
c1 = torch.from_numpy((rs.rand(100_000, 3) * [20, 20, 2]).astype(np.float32)).cuda()
run("config1 (100k, r = 5e)", c1, [0.1, 0.2, 0.4], [0.5, 1.0, 2.0], reps=10)
