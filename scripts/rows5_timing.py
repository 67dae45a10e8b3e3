"""dev tool: feature-phase time of the 11x11x11-window kernel on the config-2 scene."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import _lib, multiscale, synth
cloud = synth.urban_scene(10_000_000, seed=20, device="cuda")
lib = _lib.lib(); lib.nbr_timing_enable(1); ph = (ctypes.c_double * 8)()
for e, r in ((0.4, 2.0), (0.2, 0.8), (0.2, 1.0), (0.1, 0.5)):
    for _ in range(2):
        out = multiscale.process_single_core(cloud, cloud, [e], [r], out_dtype=np.float32)
    torch.cuda.synchronize(); lib.nbr_timing_read(ph)
    for _ in range(3):
        out = multiscale.process_single_core(cloud, cloud, [e], [r], out_dtype=np.float32)
    torch.cuda.synchronize(); lib.nbr_timing_read(ph)
    print("e=%.1f r=%.1f (r/e=%.0f): features %.3f ms  mean pop %.1f" % (e, r, r / e, ph[3] / 3, out[:, 0].mean().item()))
