"""BASELINE config 5 (scaled down): GPU multiscale features feeding the classifier hand-off give the same
confusion matrix, within sampling noise, as the oracle's features on the same labelled scene."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

EDGES = (0.1, 0.2, 0.4, 0.8, 1.6)
RADII = (0.3, 0.6, 1.2, 2.4, 4.8)


def test_scaleset_block_equals_flat_call():
    from nimrud_b200 import learning, multiscale, synth
    cloud = synth.urban_scene(40_000, seed=23).cuda()
    scaleset = [(0.2, [0.4, 0.6, 1.0]), (0.4, [1.2])]
    block = learning.scaleset_features(cloud, cloud, scaleset)
    flat = multiscale.process_single_core(cloud, cloud, [0.2, 0.2, 0.2, 0.4], [0.4, 0.6, 1.0, 1.2], out_dtype=np.float32)
    assert block.shape == (40_000, 16)
    assert torch.equal(block, torch.nan_to_num(flat))


def test_classifier_handoff_matches_oracle_features(c_oracle):
    from nimrud_b200 import learning, synth
    cloud, labels = synth.urban_scene(120_000, seed=23, return_labels=True)
    gpu = learning.classify_scene(cloud.cuda(), labels, EDGES, RADII, per_class=2000, trials=3, seed=1)
    feats = gpu["features"]
    assert feats.shape == (120_000, 20) and feats.is_cuda
    # oracle features for a labelled subset (the CPU oracle is slow): same classifier recipe on both
    rs = np.random.RandomState(0)
    sub = np.sort(rs.choice(120_000, 12_000, replace=False))
    cl = cloud.numpy()
    ref = c_oracle.process(cl[sub], cl, EDGES, RADII)
    got = feats[torch.as_tensor(sub).cuda()].cpu().numpy().astype(np.float64)
    assert np.array_equal(got[:, 0::4], ref[:, 0::4])                       # populations identical
    lab = labels.numpy()[sub]
    a = learning.classify_scene(None, lab, EDGES, RADII, per_class=400, trials=4, seed=2, feats=got)
    b = learning.classify_scene(None, lab, EDGES, RADII, per_class=400, trials=4, seed=2, feats=ref)
    # same seed, same rows, features equal to 1e-4: the forests and their confusion matrices agree
    diff = np.abs(a["confusion_mean"] - b["confusion_mean"]).max()
    assert diff <= 0.02 * a["confusion_mean"].sum(0).max(), diff
    assert gpu["producer"].mean() > 60.0                                    # the 4 components are separable


def test_scaleset_drop_empty_rule():
    """legacy row rule (prototypes/apc.py:565, 655-660): a query is represented if one pass at least gave it a feature."""
    from nimrud_b200 import learning, synth
    cloud = synth.urban_scene(20_000, seed=3).numpy()
    far = np.array([[1e4, 1e4, 50.0], [-500.0, 3.0, 2.0]], dtype=np.float32)
    q = np.concatenate([cloud[:500], far])
    scaleset = [(0.2, [0.4, 0.6]), (0.4, [1.2])]
    full = learning.scaleset_features(q, cloud, scaleset)
    idx, kept = learning.scaleset_features(q, cloud, scaleset, drop_empty=True)
    assert full.shape == (502, 12)
    assert np.array_equal(idx, np.nonzero((full[:, 0::4] > 0).any(1))[0])
    assert 500 not in idx and 501 not in idx and len(idx) >= 490
    assert np.array_equal(kept, full[idx])
