"""
the oracle against (a) the reference's own known-answer tests for the voxel filter
(nimrud/utils/tests/geometry_tests.py:84-279) and (b) golden vectors produced by running the
reference itself (tests/golden/make_golden.py).  CPU only.
"""
import numpy as np
import pytest

from conftest import load_golden
from oracle import nimrud_oracle as O

BOUNDS = np.asarray([[0, 0, 0], [100, 100, 100]])


def test_voxel_shift_known_answers():
    # geometry_tests.py:84-139
    for dim in (2, 3):
        p = O.grid_params(BOUNDS[:, :dim], 0.001)
        assert np.array_equal(p.widths, [17, 17, 17][:dim])
        assert np.array_equal(p.shifts, [17, 34][:dim - 1])
        with pytest.raises(ValueError):
            O.grid_params(BOUNDS[:, :dim], 0.00001 if dim == 3 else 0.00000001)


def test_voxel_masks_known_answers():
    # geometry_tests.py:142-158
    for dim in (2, 3):
        p = O.grid_params(BOUNDS[:, :dim], 1)
        assert np.array_equal(p.masks, [0b1111111, 0b11111110000000, 0b111111100000000000000][:dim])


def test_voxel_init_errors():
    # geometry_tests.py:17-80
    with pytest.raises(ValueError):
        O.grid_params(np.random.rand(1, 3), 0.5)
    for dim in (1, 4):
        with pytest.raises(ValueError):
            O.grid_params(np.random.rand(10, dim), 0.5)
    with pytest.raises(ValueError):
        O.grid_params(np.random.rand(10), 0.5)
    pts = np.random.RandomState(10).rand(1000, 3) * 100
    p = O.grid_params(pts, 0.5)
    assert np.array_equal(p.minimum_corner, pts.min(0) - 0.25)
    assert np.array_equal(p.maximum_corner, pts.max(0) + 0.25)


def test_voxel_address_known_answer():
    # geometry_tests.py:196-256: (10,11,12), e=1 -> 198026 and back
    p = O.grid_params(BOUNDS, 1)
    assert O.coordinate_to_address(p, np.arange(3) + 10)[0] == 198026
    assert np.allclose(O.address_to_coordinate(p, 198026).flatten(), np.arange(3) + 10)
    with pytest.raises(ValueError):
        O.coordinate_to_address(p, np.zeros((1, 3)) - 1.5)
    with pytest.raises(ValueError):
        O.coordinate_to_address(p, np.zeros((1, 3)) + 101.5)
    O.coordinate_to_address(p, np.zeros((1, 3)) + 100.5)


def test_voxel_unique_known_answer():
    # geometry_tests.py:261-279
    for dim in (2, 3):
        p = O.grid_params(BOUNDS[:, :dim], 1)
        pts = np.concatenate([np.zeros((1, dim)) + off for off in np.arange(0, 20, 2)])
        _, centres = O.unique_voxels(p, np.vstack((pts, pts)))
        assert np.array_equal(pts, centres)


@pytest.mark.parametrize("name", ["small", "urban", "degenerate"])
def test_python_oracle_matches_reference_bitwise(name):
    g = load_golden(name)
    q = g["query"].astype(np.float64)
    s = g["search"].astype(np.float64)
    out = O.process(q, s, g["edges"], g["radii"])
    assert np.array_equal(out, g["features"])
    for i, (e, r) in enumerate(zip(g["edges"], g["radii"])):
        p = O.grid_params(s, e)
        assert np.array_equal(p.minimum_corner, g["s%d_min_corner" % i])
        assert np.array_equal(p.widths, g["s%d_widths" % i])
        keys, centres = O.unique_voxels(p, s)
        assert np.array_equal(keys, g["s%d_addresses" % i])
        assert np.array_equal(centres, g["s%d_centres" % i])
        off, idx = O.radius_sets(q, centres, r)
        assert np.array_equal(off, g["s%d_offsets" % i])
        assert np.array_equal(idx, g["s%d_indices" % i])


def test_membership_predicate_is_the_literal_expression():
    # scipy's tree query == dx*dx+dy*dy+dz*dz <= r*r in float64, inclusive
    g = load_golden("degenerate")
    q = g["query"].astype(np.float64)
    s = g["search"].astype(np.float64)
    for e, r in zip(g["edges"], g["radii"]):
        _, centres = O.unique_voxels(O.grid_params(s, e), s)
        a = O.radius_sets(q, centres, r)
        b = O.radius_sets_bruteforce(q, centres, r)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # an exactly-on-the-boundary neighbor is a member
    centres = np.array([[0.0, 0.0, 0.0], [3.0, 0.0, 0.0], [3.0000000000000004, 0.0, 0.0]])
    off, idx = O.radius_sets(np.zeros((1, 3)), centres, 3.0)
    assert list(idx) == [0, 1]


def test_python_oracle_config1_subset():
    g = load_golden("config1")
    cloud = (np.random.RandomState(10).rand(100_000, 3) * [20, 20, 2]).astype(np.float32).astype(np.float64)
    pick = g["pick"][:60]
    out = O.process(cloud[pick], cloud, g["edges"], g["radii"])
    assert np.array_equal(out, g["features"][:60])


@pytest.mark.parametrize("name", ["small", "urban", "degenerate"])
def test_c_oracle_matches_reference(name, c_oracle):
    g = load_golden(name)
    q = g["query"].astype(np.float64)
    s = g["search"].astype(np.float64)
    out = c_oracle.process(q, s, g["edges"], g["radii"])
    ref = g["features"]
    assert np.array_equal(out[:, 0::4], ref[:, 0::4])                  # populations exact
    assert np.abs(out - ref).max() < 1e-12                            # float64 rounding only
    for i, (e, r) in enumerate(zip(g["edges"], g["radii"])):
        minc, maxc, widths = c_oracle.grid_widths(s, e)
        assert np.array_equal(minc, g["s%d_min_corner" % i])
        assert np.array_equal(widths, g["s%d_widths" % i])
        keys, centres = c_oracle.unique_voxels(s, minc, e, widths)
        assert np.array_equal(keys, g["s%d_addresses" % i])
        assert np.array_equal(centres, g["s%d_centres" % i])
        off, idx = c_oracle.radius_sets(q, keys, minc, e, widths, r)
        assert np.array_equal(off, g["s%d_offsets" % i])
        assert np.array_equal(idx, g["s%d_indices" % i])


def test_c_oracle_config1(c_oracle):
    g = load_golden("config1")
    cloud = (np.random.RandomState(10).rand(100_000, 3) * [20, 20, 2]).astype(np.float32).astype(np.float64)
    out = c_oracle.process(cloud[g["pick"]], cloud, g["edges"], g["radii"])
    assert np.array_equal(out[:, 0::4], g["features"][:, 0::4])
    assert np.abs(out - g["features"]).max() < 1e-12
    minc, maxc, widths = c_oracle.grid_widths(cloud, 0.1)
    keys, _ = c_oracle.unique_voxels(cloud, minc, 0.1, widths)
    assert len(keys) == g["n_voxels"][0] == 94187


def test_knn_oracles_agree(c_oracle):
    rs = np.random.RandomState(3)
    pts = np.round(rs.rand(3000, 3) * 10) * 0.5           # lattice: many exact ties
    q = np.concatenate([pts[:40], rs.rand(40, 3) * 5])
    ia, da = O.knn_sets(q, pts, 12)
    ib, db = c_oracle.knn(q, pts, 12)
    assert np.array_equal(ia, ib) and np.array_equal(da, db)
    fa = O.knn_features(q, pts, [4, 12])
    fb = c_oracle.knn_features(q, pts, ib, [4, 12])
    assert np.abs(fa - fb).max() < 1e-12


def test_sharding_invariance(c_oracle):
    g = load_golden("small")
    q = g["query"].astype(np.float64)
    whole = O.process(q[:200], q, g["edges"][:1], g["radii"][:1])
    parts = np.concatenate([O.process(q[a:a + 50], q, g["edges"][:1], g["radii"][:1]) for a in range(0, 200, 50)])
    assert np.array_equal(whole, parts)


def test_extended_row_is_self_consistent():
    """the extension descriptors have no reference code (parity unpinned): check the oracle against the
    definitions -- eigen-decomposition identities and the sign conventions of the emitted vectors."""
    from oracle import nimrud_oracle as O
    rs = np.random.RandomState(3)
    for _ in range(50):
        nb = rs.rand(rs.randint(4, 40), 3) * [1.0, 0.6, 0.2]
        row = O.extended_row(nb)
        assert row.shape == (22,)
        cov = np.cov(nb, rowvar=False)
        w, v = np.linalg.eigh(cov)
        e = w[::-1] / w.sum()
        assert np.allclose(row[:3], [(e[0] - e[1]) / e[0], (e[1] - e[2]) / e[0], e[2] / e[0]], rtol=1e-9, atol=1e-12)
        assert np.isclose(row[11], np.trace(cov))
        assert np.allclose(row[12:18], cov[np.triu_indices(3)])
        normal = row[8:11]
        assert normal[2] >= 0 and np.isclose(np.linalg.norm(normal), 1.0)
        assert np.allclose(cov @ normal, w[0] * normal, atol=1e-9)
        # x, y of the eigenvectors of the largest / middle eigenvalue, sign x > 0
        assert row[18] >= 0 and row[20] >= 0
        for cols, k in ((row[18:20], 2), (row[20:22], 1)):
            full = v[:, k] if v[0, k] >= 0 else -v[:, k]
            assert np.allclose(cols, full[:2], atol=1e-12)
    assert not O.extended_row(np.zeros((2, 3))).any()          # undefined -> zeros
