"""dev tool: time the fused feature kernel one scale at a time on the config-2 scene."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import _lib, multiscale, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cloud = synth.urban_scene(n, seed=20, device="cuda")
lib = _lib.lib()
lib.nbr_timing_enable(1)
ph = (ctypes.c_double * 8)()
for e, r in ((0.1, 0.3), (0.2, 0.6), (0.4, 1.2), (0.8, 2.4), (1.6, 4.8)):
    for _ in range(2):
        out = multiscale.process_single_core(cloud, cloud, [e], [r], out_dtype=np.float32)
    torch.cuda.synchronize(); lib.nbr_timing_read(ph)
    for _ in range(3):
        out = multiscale.process_single_core(cloud, cloud, [e], [r], out_dtype=np.float32)
    torch.cuda.synchronize(); lib.nbr_timing_read(ph)
    print("e=%.1f r=%.1f  features %.3f ms  index %.3f  order %.3f  mean pop %.1f" % (e, r, ph[3] / 3, ph[1] / 3, ph[2] / 3, out[:, 0].mean().item()))
