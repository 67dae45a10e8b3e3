"""dev tool: one process driving two devices in turn (kernel attributes, table caches and scratch are per device)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import multiscale, synth
cloud = synth.urban_scene(60_000, seed=5, device="cpu")
order = (1, 0, 1) if len(sys.argv) > 1 else (0, 1, 0, 1)
for edges, radii in (((0.2,), (0.6,)), ((0.2,), (1.0,)), ((0.2,), (1.4,))):
    outs = []
    for d in order:
        c = cloud.to("cuda:%d" % d)
        print("radius", radii, "device", d, flush=True)
        outs.append(multiscale.process_single_core(c, c, edges, radii, out_dtype=np.float32).cpu())
    print(radii, [bool(torch.equal(outs[0], o)) for o in outs], [(outs[0] != o).sum().item() for o in outs], flush=True)
knns = []
for d in order:
    c = cloud.to("cuda:%d" % d)
    print("knn device", d, flush=True)
    knns.append(multiscale.knn_features(c, c, 0.2, (5, 10), out_dtype=np.float32).cpu())
print("knn", [bool(torch.equal(knns[0], o)) for o in knns], [(knns[0] != o).sum().item() for o in knns])
