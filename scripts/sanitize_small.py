"""dev tool: one small pass through every kernel family, for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import multiscale, synth, learning
cloud = synth.urban_scene(30_000, seed=9, device="cuda")
q = cloud[::3].contiguous()
# rows3 (shell tables, staged + direct windows, row buffer), interval kernel (r/e = 5), exact kernel (r/e = 12)
a = multiscale.process_single_core(cloud, cloud, [0.1, 0.2, 0.4, 0.8, 1.6], [0.3, 0.6, 1.2, 2.4, 4.8], out_dtype=np.float32)
b = multiscale.process_single_core(q, cloud, [0.2, 0.2, 0.2, 0.1], [0.4, 0.6, 1.0, 1.2])
c = multiscale.process_single_core(q, cloud, [0.2], [0.6], descriptors="extended")
# tile + halo entry (queries = prefix of the search buffer)
d = multiscale.process_single_core(cloud[:20_000], cloud, [0.2, 0.4], [0.6, 1.2], out_dtype=np.float32)
index = multiscale.LatticeIndex(cloud, 0.2, indexed=True)
off, idx = index.radius_sets(q, 0.6)
i10, d10, f10 = index.knn(q, 50, ks=(10, 20, 50), out_dtype=np.float32)
index.close()
# 11-wide windows (r/e = 4, 5), vector field operator
e = multiscale.process_single_core(q, cloud, [0.2, 0.1], [0.8, 0.5], out_dtype=np.float32)
v = multiscale.vector_field_features(q, cloud, torch.rand(cloud.shape[0], 3, device="cuda"), 0.2, [0.6, 1.0])
# multi-GPU tile path on one device: halo mailboxes (box table, push, wait), tile order, lattices from tile + mailbox
from nimrud_b200 import distributed as nd
xs = cloud[:, 0]
cut = [xs.quantile(0.33).item(), xs.quantile(0.66).item()]
tiles = [cloud[xs < cut[0]].contiguous(), cloud[(xs >= cut[0]) & (xs < cut[1])].contiguous(), cloud[xs >= cut[1]].contiguous()]
t = nd.process_tiles_local(tiles, [0.1, 0.2, 0.4], [0.3, 0.6, 1.2], capacity_rows=40_000)
# host-buffer path (pinned rings, pieces, host threads)
h = multiscale.process_single_core(cloud.cpu().numpy(), cloud.cpu().numpy(), [0.2, 0.4], [0.6, 1.2])
torch.cuda.synchronize()
print("ok", a.shape, b.shape, c.shape, d.shape, int(off[-1]), i10.shape, e.shape, v.shape, [x.shape for x in t], h.shape)
