import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from nimrud_b200 import multiscale, synth
cloud = synth.urban_scene(10_000_000, seed=20, device="cuda")
for _ in range(3):
    out = multiscale.process_single_core(cloud, cloud, [0.4], [2.0], out_dtype=np.float32)
torch.cuda.synchronize()
