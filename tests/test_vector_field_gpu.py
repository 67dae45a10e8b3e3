"""vector-field multiscale operator (extension, parity unpinned) against a numpy restatement built on the
oracle's voxel filter and radius sets."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference(query, search, vectors, edge, radii, c_oracle):
    s = search.astype(np.float64)
    minc, maxc, widths = c_oracle.grid_widths(s, edge)
    ukeys, centres = c_oracle.unique_voxels(s, minc, edge, widths)
    # voxel of every search point: position of its centre among the unique centres (np.unique order)
    cells = np.floor((s - minc) / edge).astype(np.int64)
    shifts = np.concatenate([[0], np.cumsum(widths)[:-1]])
    keys = (cells[:, 0] << shifts[0]) + (cells[:, 1] << shifts[1]) + (cells[:, 2] << shifts[2])
    rank = np.searchsorted(ukeys, keys)
    assert np.array_equal(ukeys[rank], keys)
    F = vectors.shape[1]
    sums = np.zeros((len(ukeys), F)); np.add.at(sums, rank, vectors.astype(np.float64))
    voxvec = (sums / np.bincount(rank, minlength=len(ukeys))[:, None]).astype(np.float32).astype(np.float64)
    out = np.zeros((len(query), F * len(radii)))
    for k, r in enumerate(radii):
        off, idx = c_oracle.radius_sets(query.astype(np.float64), ukeys, minc, edge, widths, r)
        for i in range(len(query)):
            members = idx[off[i]:off[i + 1]]
            if len(members):
                out[i, k * F:(k + 1) * F] = voxvec[members].mean(0)
    return out


@pytest.mark.parametrize("F", [1, 3, 11])
def test_vector_field_means(F, c_oracle):
    from nimrud_b200 import multiscale, synth
    rs = np.random.RandomState(F)
    cloud = synth.urban_scene(40_000, seed=14).numpy()
    vectors = rs.randn(len(cloud), F).astype(np.float32)
    q = np.concatenate([cloud[::40], cloud[:5] + 100.0]).astype(np.float32)       # the last 5: empty neighborhoods
    radii = (0.3, 0.9)
    got = multiscale.vector_field_features(q, cloud, vectors, 0.2, radii)
    ref = _reference(q, cloud, vectors, 0.2, radii, c_oracle)
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    assert np.all(got[-5:] == 0)
