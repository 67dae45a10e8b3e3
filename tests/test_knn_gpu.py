"""GPU kNN against the brute-force oracle: exact index sets in (d^2, index) order, ties included."""
import numpy as np
import pytest
import torch

from conftest import assert_features_close

pytestmark = pytest.mark.gpu


def _centres(search, edge):
    from oracle import nimrud_oracle as O
    s = search.astype(np.float64)
    return O.unique_voxels(O.grid_params(s, edge), s)[1]


@pytest.mark.parametrize("k", [1, 10, 50, 128])
def test_knn_sets_bit_exact_with_ties(k, c_oracle):
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(60_000, seed=21)
    q = synth.with_ties(cloud[:3000], 0.1, seed=21, fraction=0.05).numpy()
    cloud = cloud.numpy()
    index = multiscale.LatticeIndex(cloud, 0.1, indexed=True)
    idx, d2 = index.knn(q, k)
    centres = _centres(cloud, 0.1)
    ref_idx, ref_d2 = c_oracle.knn(q.astype(np.float64), centres, k)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(d2.cpu().numpy(), ref_d2)
    index.close()


def test_knn_fewer_points_than_k_and_far_queries(c_oracle):
    from nimrud_b200 import multiscale
    rs = np.random.RandomState(2)
    cloud = (rs.rand(30, 3) * 2).astype(np.float32)
    q = np.concatenate([cloud[:5], [[50.0, -40.0, 3.0]], [[-1000.0, 0.0, 0.0]]]).astype(np.float32)
    index = multiscale.LatticeIndex(cloud, 0.25, indexed=True)
    idx, d2 = index.knn(q, 40)
    centres = _centres(cloud, 0.25)
    ref_idx, ref_d2 = c_oracle.knn(q.astype(np.float64), centres, 40)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(d2.cpu().numpy(), ref_d2)
    index.close()


def test_knn_multiscale_features(c_oracle):
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(80_000, seed=22).numpy()
    q = cloud[:4000]
    ks = (10, 20, 50)
    feats = multiscale.knn_features(q, cloud, 0.1, ks)
    centres = _centres(cloud, 0.1)
    ref_idx, _ = c_oracle.knn(q.astype(np.float64), centres, 50)
    ref = c_oracle.knn_features(q.astype(np.float64), centres, ref_idx, ks)
    assert feats.shape == (4000, 12)
    # centroid tolerance is relative to the k-th distance scale; use the largest k-th distance
    assert_features_close(feats, ref, [1.0, 1.0, 1.0])


@pytest.mark.parametrize("k", [1, 10, 50, 128])
def test_knn_over_raw_points_bit_exact_with_ties(k, c_oracle):
    """search over the unfiltered cloud (legacy sspedge = 0): exact (d^2, index) order against the brute force,
    with duplicated search points (equal distances, different indices) and lattice-aligned queries."""
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(40_000, seed=23)
    search = torch.cat([cloud, cloud[:2000]], 0).numpy()              # 2000 exact duplicates in the search set
    q = synth.with_ties(cloud[:2500], 0.1, seed=21, fraction=0.05).numpy()
    idx, d2 = multiscale.knn_points(q, search, k)
    ref_idx, ref_d2 = c_oracle.knn(q.astype(np.float64), search.astype(np.float64), k)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(d2.cpu().numpy(), ref_d2)


def test_knn_over_raw_points_small_far_and_features(c_oracle):
    from nimrud_b200 import multiscale, synth
    rs = np.random.RandomState(5)
    cloud = (rs.rand(30, 3) * 2).astype(np.float32)
    q = np.concatenate([cloud[:5], [[50.0, -40.0, 3.0]], [[-1000.0, 0.0, 0.0]]]).astype(np.float32)
    idx, d2 = multiscale.knn_points(q, cloud, 40)                      # fewer points than k: padded with -1 / inf
    ref_idx, ref_d2 = c_oracle.knn(q.astype(np.float64), cloud.astype(np.float64), 40)
    assert np.array_equal(idx.cpu().numpy(), ref_idx)
    assert np.array_equal(d2.cpu().numpy(), ref_d2)
    scene = synth.urban_scene(60_000, seed=24).numpy()
    qs = scene[:3000]
    ks = (10, 20, 50)
    feats = multiscale.knn_features(qs, scene, 0, ks)                  # edge 0: raw points
    ref_idx, _ = c_oracle.knn(qs.astype(np.float64), scene.astype(np.float64), 50)
    ref = c_oracle.knn_features(qs.astype(np.float64), scene.astype(np.float64), ref_idx, ks)
    assert feats.shape == (3000, 12)
    assert_features_close(feats, ref, [1.0, 1.0, 1.0])
