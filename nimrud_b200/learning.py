"""
classifier hand-off: what consumes the multiscale feature block downstream.

`nimrud/learning` is an empty package in the reference; the only description of the consumer is the
scripting in nimrud/prototypes/apc.py (gmso_APC :497-680, multiclass_self/blind :807-1235,
balance_resampler :1576-1637) and the metrics in nimrud/prototypes/ml.py (:465-555).  this module keeps
that contract without the interactive input() prompts and pickled archives:

  scaleset_features      the caller contract of gmso_APC: scaleset = [(voxel_edge, [radii...]), ...] ->
                         one dense (N, 4 * total_scales) block, scale-major, undefined features = 0
                         (apc.py:514-518, 640-663; pull_feats nan_to_num :313-318)
  balanced_training_set  per-class balanced sampling, size = smallest class (apc.py:1143-1162)
  fit_classifier         ExtraTreesClassifier(n_jobs=4, n_estimators=30, gini, no bootstrap) (apc.py:1191)
  balance_resampler      repeated balanced validation -> mean / std confusion matrix (apc.py:1576-1637)
  mc_confusion, user_producer, three_metrics      (ml.py:521, :465, :491)

features are computed on the GPU (nimrud_b200.multiscale) and stay there until a sample of them is
needed on the host: only the balanced sample crosses PCIe for training.  the classifier itself is
scikit-learn on the host, as in the reference.
"""
import numpy as np
import torch

from . import multiscale
from ._util import is_torch


# --------------------------------------------------------------------------------------------------
# feature block
# --------------------------------------------------------------------------------------------------
def scaleset_features(query_cloud, search_cloud, scaleset, out_dtype=np.float32, descriptors="reference",
                      drop_empty=False):
    """
    scaleset = [(voxel_edge, [radius, ...]), ...]  ->  (N, C * total_scales) features, scale-major in the
    caller's order (C = 4 reference columns, or 26 with descriptors="extended").  every radius of a group
    shares the group's voxel lattice: one index build and one staged window per group on the device.

    drop_empty=True follows the legacy driver's row rule (nimrud/prototypes/apc.py:565, 655-660: "if a point gets
    a feature from at least one pass ... it will be represented in the index set"): returns (index, features) with
    only the queries whose neighborhood is non-empty at one scale at least, index = their rows in query_cloud.
    """
    edges, radii = [], []
    for edge, group in scaleset:
        for r in group:
            edges.append(float(edge))
            radii.append(float(r))
    feats = multiscale.process_single_core(query_cloud, search_cloud, edges, radii, out_dtype=out_dtype,
                                           descriptors=descriptors)
    feats = torch.nan_to_num(feats) if is_torch(feats) else np.nan_to_num(feats)
    if not drop_empty:
        return feats
    ncol = feats.shape[1] // max(len(radii), 1) if len(radii) else 4
    populated = (feats[:, 0::ncol] > 0).any(1) if len(radii) else feats[:, :0].any(1)
    index = populated.nonzero()[:, 0] if is_torch(feats) else np.nonzero(populated)[0]
    return index, feats[index]


# --------------------------------------------------------------------------------------------------
# sampling / classifier
# --------------------------------------------------------------------------------------------------
def _label_sets(labels):
    labels = np.asarray(labels).astype(np.int64).ravel()
    numlabs = int(labels.max()) + 1
    idx = np.arange(labels.size)
    return [idx[labels == n] for n in range(numlabs)]


def balanced_indices(labels, per_class=None, rng=None):
    """row indices and labels of a class-balanced sample: `per_class` rows of every class (default: the
    population of the smallest class), drawn without replacement."""
    rng = rng if rng is not None else np.random
    labset = _label_sets(labels)
    smallest = min(s.size for s in labset)
    tnum = smallest if per_class is None else min(int(per_class), smallest)
    rows, labs = [], []
    for n, members in enumerate(labset):
        rows.append(rng.permutation(members)[:tnum])
        labs.append(np.full(tnum, n, dtype=np.int64))
    return np.concatenate(rows), np.concatenate(labs)


def take_rows(feats, rows):
    """feats[rows] as a host float array; a CUDA feature block is gathered on the device first."""
    if is_torch(feats):
        sel = torch.as_tensor(rows, device=feats.device, dtype=torch.long)
        return feats.index_select(0, sel).cpu().numpy()
    return np.asarray(feats).take(rows, axis=0)


def balanced_training_set(feats, labels, per_class=None, rng=None):
    rows, labs = balanced_indices(labels, per_class, rng)
    return take_rows(feats, rows), labs


def fit_classifier(tset, tlabels, n_jobs=4, n_estimators=30, random_state=None):
    from sklearn.ensemble import ExtraTreesClassifier
    clf = ExtraTreesClassifier(n_jobs=n_jobs, n_estimators=n_estimators, criterion="gini", bootstrap=False,
                               random_state=random_state)
    clf.fit(tset, tlabels)
    return clf


# --------------------------------------------------------------------------------------------------
# metrics
# --------------------------------------------------------------------------------------------------
def mc_confusion(lies, truth):
    """conf[row, col] = number of points of known class `col` that were assigned class `row`."""
    lies = np.asarray(lies).astype(np.int64).ravel()
    truth = np.asarray(truth).astype(np.int64).ravel()
    nlabels = int(max(truth.max(), lies.max())) + 1
    conf = np.bincount(lies * nlabels + truth, minlength=nlabels * nlabels).reshape(nlabels, nlabels)
    return conf.astype(np.float64)


def user_producer(conf):
    """user (per assigned class, over rows) and producer (per known class, over columns) accuracy in %."""
    conf = np.asarray(conf, dtype=np.float64)
    diag = np.diag(conf)
    with np.errstate(divide="ignore", invalid="ignore"):
        return diag / conf.sum(1) * 100, diag / conf.sum(0) * 100


def three_metrics(conf):
    """per class: true positive, false positive, false negative rates (known classes assumed balanced)."""
    conf = np.asarray(conf, dtype=np.float64)
    diag = np.diag(conf)
    n_real = conf.sum(0)[0]
    n_pred = conf.sum(1)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.column_stack((diag / n_real, (n_real - diag) / n_real, (n_pred - diag) / n_pred))


def balance_resampler(feats, labels, clf, trials, rng=None):
    """mean and std confusion matrix over `trials` balanced validation samples of half the smallest class."""
    rng = rng if rng is not None else np.random
    labset = _label_sets(labels)
    numlabs = len(labset)
    vnum = int(np.floor(0.5 * min(s.size for s in labset)))
    cmat = np.zeros((numlabs, numlabs, trials))
    for t in range(trials):
        rows, vlabels = balanced_indices(labels, vnum, rng)
        assigned = clf.predict(take_rows(feats, rows))
        cmat[:, :, t] = _pad(mc_confusion(assigned, vlabels), numlabs)
    return cmat.mean(2), cmat.std(2)


def _pad(conf, n):
    out = np.zeros((n, n))
    out[:conf.shape[0], :conf.shape[1]] = conf[:n, :n]
    return out


# --------------------------------------------------------------------------------------------------
# the whole hand-off (BASELINE config 5)
# --------------------------------------------------------------------------------------------------
def classify_scene(cloud, labels, edge_lengths, radii, per_class=None, trials=3, seed=0, feats=None):
    """
    labelled scene -> GPU multiscale features -> balanced ExtraTrees -> balanced validation.
    returns dict(features, classifier, confusion_mean, confusion_std, user, producer).
    `feats` may be supplied to classify with an existing feature block.
    """
    rng = np.random.RandomState(seed)
    if feats is None:
        feats = multiscale.process_single_core(cloud, cloud, list(edge_lengths), list(radii), out_dtype=np.float32)
        feats = torch.nan_to_num(feats) if is_torch(feats) else np.nan_to_num(feats)
    labels = labels.cpu().numpy() if is_torch(labels) else np.asarray(labels)
    tset, tlabels = balanced_training_set(feats, labels, per_class, rng)
    clf = fit_classifier(tset, tlabels, random_state=seed)
    mean, std = balance_resampler(feats, labels, clf, trials, rng)
    user, prod = user_producer(mean)
    return {"features": feats, "classifier": clf, "confusion_mean": mean, "confusion_std": std, "user": user,
            "producer": prod}
