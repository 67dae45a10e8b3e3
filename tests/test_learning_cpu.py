"""classifier hand-off metrics against golden vectors produced by the reference's own functions
(nimrud/prototypes/ml.py:465-555, see tests/golden/make_golden_learning.py), and the sampling logic."""
import numpy as np

from conftest import load_golden


def test_metrics_match_reference_golden():
    from nimrud_b200 import learning
    g = load_golden("learning")
    for case in range(3):
        conf = learning.mc_confusion(g["lies%d" % case], g["truth%d" % case])
        assert np.array_equal(conf, g["conf%d" % case])
        user, prod = learning.user_producer(conf)
        assert np.allclose(user, g["user%d" % case], rtol=1e-14, atol=0)
        assert np.allclose(prod, g["prod%d" % case], rtol=1e-14, atol=0)
        assert np.allclose(learning.three_metrics(conf), g["three%d" % case], rtol=1e-14, atol=0)


def test_balanced_sampling():
    from nimrud_b200 import learning
    rs = np.random.RandomState(3)
    labels = np.concatenate([np.zeros(500), np.ones(120), np.full(300, 2)]).astype(np.int64)
    rs.shuffle(labels)
    feats = np.arange(labels.size, dtype=np.float64)[:, None] * np.ones((1, 3))
    rows, labs = learning.balanced_indices(labels, rng=np.random.RandomState(1))
    assert rows.size == 3 * 120 and np.unique(rows).size == rows.size          # smallest class, no repeats
    assert np.array_equal(labels[rows], labs)
    assert np.array_equal(np.bincount(labs), [120, 120, 120])
    tset, tl = learning.balanced_training_set(feats, labels, per_class=50, rng=np.random.RandomState(1))
    assert tset.shape == (150, 3) and np.array_equal(labels[tset[:, 0].astype(int)], tl)


def test_classifier_and_resampler_on_separable_blobs():
    from nimrud_b200 import learning
    rs = np.random.RandomState(5)
    centres = np.array([[0, 0, 0, 0], [4, 0, 0, 0], [0, 4, 0, 0]], dtype=np.float64)
    labels = rs.randint(0, 3, 3000)
    feats = centres[labels] + rs.randn(3000, 4) * 0.3
    tset, tl = learning.balanced_training_set(feats, labels, per_class=200, rng=rs)
    clf = learning.fit_classifier(tset, tl, random_state=0)
    mean, std = learning.balance_resampler(feats, labels, clf, trials=2, rng=rs)
    assert mean.shape == (3, 3) and std.shape == (3, 3)
    user, prod = learning.user_producer(mean)
    assert user.min() > 99 and prod.min() > 99
