"""BASELINE config 5: GPU multiscale features feeding the classifier hand-off on a synthetic labelled 20M-point scene.
features (N, 20) on the GPU -> balanced sample -> ExtraTrees(30, gini) -> balanced validation (nimrud_b200.learning);
then the same recipe on oracle features of a 30k-row subset, confusion matrices side by side."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nimrud_b200 import learning, multiscale, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
EDGES = (0.1, 0.2, 0.4, 0.8, 1.6); RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
cloud, labels = synth.urban_scene(N, seed=23, device="cuda", return_labels=True)
torch.cuda.synchronize()
multiscale.process_single_core(cloud, cloud, EDGES, RADII, out_dtype=np.float32); torch.cuda.synchronize()
t0 = time.perf_counter()
feats = multiscale.process_single_core(cloud, cloud, EDGES, RADII, out_dtype=np.float32)
torch.cuda.synchronize(); t1 = time.perf_counter()
print("features: %d points x 5 scales in %.1f ms (%.2f G point*scales/s)" % (N, (t1 - t0) * 1e3, N * 5 / (t1 - t0) / 1e9))
res = learning.classify_scene(cloud, labels, EDGES, RADII, per_class=20000, trials=3, seed=1, feats=torch.nan_to_num(feats))
t2 = time.perf_counter()
print("classifier hand-off (balanced sample of 4 x 20000 rows, ExtraTrees 30, 3 validation trials): %.1f s" % (t2 - t1))
np.set_printoptions(precision=1, suppress=True)
print("confusion (mean of 3 balanced validations, rows = assigned, cols = known: ground, building, pole/wire, vegetation):")
print(res["confusion_mean"])
print("user %s producer %s" % (res["user"], res["producer"]))
# oracle features for a subset, same recipe on both feature sets
from oracle import c_oracle
c_oracle.build()
rs = np.random.RandomState(0)
cl = cloud.cpu().numpy()
centre = cl[:, :2].mean(0)
near = np.abs(cl[:, :2] - centre).max(1) < 45.0
sub_idx = np.nonzero(np.abs(cl[:, :2] - centre).max(1) < 35.0)[0]
sub_idx = np.sort(rs.choice(sub_idx, min(30000, len(sub_idx)), replace=False))
# the oracle anchors its grids on the cloud it is given: hand it the WHOLE cloud's corner points too
corners = np.stack([cl.min(0), cl.max(0)]).astype(np.float64)
search = np.concatenate([cl[near].astype(np.float64), corners])
ref = c_oracle.process(cl[sub_idx].astype(np.float64), search, EDGES, RADII, threads=16)
got = feats[torch.from_numpy(sub_idx).cuda()].cpu().numpy().astype(np.float64)
print("subset of %d rows: populations identical: %s, max |ratio diff| %.2e" % (
    len(sub_idx), bool(np.array_equal(got[:, 0::4], ref[:, 0::4])), np.abs(got[:, 2::4] - ref[:, 2::4]).max()))
lab = labels.cpu().numpy()[sub_idx]
a = learning.classify_scene(None, lab, EDGES, RADII, per_class=None, trials=4, seed=2, feats=got)
b = learning.classify_scene(None, lab, EDGES, RADII, per_class=None, trials=4, seed=2, feats=ref)
print("confusion from GPU features:\n%s\nconfusion from oracle features:\n%s" % (a["confusion_mean"], b["confusion_mean"]))
