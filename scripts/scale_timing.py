"""dev tool: feature-phase time of single (edge, radius) pairs on the config-2 scene."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import _lib, multiscale, synth
lib = _lib.lib(); lib.nbr_timing_enable(1); ph = (ctypes.c_double * 8)()
cloud = synth.urban_scene(int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, seed=20, device="cuda")
for e, r in ((0.2, 0.6), (0.2, 0.8), (0.2, 1.0), (0.2, 1.2), (0.4, 2.0), (1.6, 8.0)):
    for _ in range(2):
        out = multiscale.process_single_core(cloud, cloud, [e], [r], out_dtype=np.float32)
    torch.cuda.synchronize(); lib.nbr_timing_read(ph)
    for _ in range(3):
        out = multiscale.process_single_core(cloud, cloud, [e], [r], out_dtype=np.float32)
    torch.cuda.synchronize(); lib.nbr_timing_read(ph)
    print("e=%.1f r=%.1f (r/e=%.1f) features %.3f ms  mean pop %.1f" % (e, r, r / e, ph[3] / 3, out[:, 0].mean().item()))
