import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from nimrud_b200 import multiscale, synth
cloud = synth.urban_scene(10_000_000, seed=20, device="cuda")
q = cloud[::5].contiguous()
for e, radii in ((0.1, (0.3,)), (0.4, (1.2, 2.0)), (1.6, (4.8, 8.0))):
    index = multiscale.LatticeIndex(cloud, e)
    fast = index.radius_features(q, radii, out_dtype=np.float64, algorithm=0)
    slow = index.radius_features(q, radii, out_dtype=np.float64, algorithm=1)
    index.close()
    for k, r in enumerate(radii):
        for c in (2, 3):
            f, s = fast[:, 4*k+c], slow[:, 4*k+c]
            d = (f - s).abs()
            ratio = d / (1e-4 * s.abs() + 1e-9)
            i = int(torch.argmax(d)); j = int(torch.argmax(ratio))
            print("e=%g r=%g col %d: max abs diff %.3e (slow %.6e, pop %d) ; max diff/tol %.3e (slow %.6e fast %.6e pop %d) ; n>1e-9: %d" % (
                e, r, c, d[i].item(), s[i].item(), int(slow[i, 4*k]), ratio[j].item(), s[j].item(), f[j].item(), int(slow[j, 4*k]), int((d > 1e-9).sum())))
