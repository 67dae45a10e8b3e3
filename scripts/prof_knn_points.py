import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import multiscale, synth
cloud = synth.urban_scene(2_000_000, seed=21, device="cuda")
for _ in range(2):
    r = multiscale.knn_points(cloud, cloud, 50, ks=(10, 20, 50), out_dtype=np.float32)
torch.cuda.synchronize()
