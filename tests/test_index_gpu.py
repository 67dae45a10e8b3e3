"""
GPU parity of the spatial index: sort / scan / unique primitives bit-exact against numpy, and the
VoxelFilter mirror against the reference's known-answer tests and the golden vectors.
"""
import ctypes

import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _lib():
    from nimrud_b200 import _lib
    return _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("n", [0, 1, 5, 255, 256, 2049, 100_003, 3_000_001])
def test_exclusive_scan_u32(n):
    L = _lib()
    rs = np.random.RandomState(n)
    a = rs.randint(0, 1000, size=n).astype(np.uint32)
    d = torch.from_numpy(a.astype(np.int64)).to("cuda").to(torch.int32)     # same bits as uint32
    out = torch.empty_like(d)
    L.check(L.lib().nbr_exclusive_scan_u32(_p(d), _p(out), n, _stream()))
    ref = np.concatenate(([0], np.cumsum(a[:-1], dtype=np.uint64))).astype(np.uint32) if n else a
    assert np.array_equal(out.cpu().numpy().view(np.uint32), ref)


@pytest.mark.parametrize("n", [1, 1000, 70_001])
def test_exclusive_scan_i64(n):
    L = _lib()
    a = np.random.RandomState(n).randint(0, 1 << 40, size=n).astype(np.int64)
    d = torch.from_numpy(a).cuda()
    out = torch.empty_like(d)
    L.check(L.lib().nbr_exclusive_scan_i64(_p(d), _p(out), n, _stream()))
    assert np.array_equal(out.cpu().numpy(), np.concatenate(([0], np.cumsum(a[:-1]))))


@pytest.mark.parametrize("n,bits", [(0, 64), (1, 64), (33, 64), (4096, 64), (4097, 20), (1_000_003, 64),
                                     (2_500_000, 35), (300_000, 8), (300_000, 3)])
def test_radix_sort_keys_bit_exact(n, bits):
    L = _lib()
    rs = np.random.RandomState(n + bits)
    a = rs.randint(0, 1 << 62, size=n, dtype=np.int64).astype(np.uint64)
    if bits < 64:
        a &= np.uint64((1 << bits) - 1)
    if n > 10:
        a[: n // 3] = a[n // 3: 2 * (n // 3)]            # duplicates
    d = torch.from_numpy(a.view(np.int64)).cuda()
    tmp = torch.empty_like(d)
    L.check(L.lib().nbr_sort_u64(_p(d), _p(tmp), n, 0, bits, _stream()))
    assert np.array_equal(d.cpu().numpy().view(np.uint64), np.sort(a))


def test_radix_sort_high_bit_keys():
    L = _lib()
    a = np.random.RandomState(1).randint(0, 1 << 63, size=50_000, dtype=np.int64).astype(np.uint64) << np.uint64(1)
    a[::7] |= np.uint64(1 << 63)
    d = torch.from_numpy(a.view(np.int64)).cuda()
    tmp = torch.empty_like(d)
    L.check(L.lib().nbr_sort_u64(_p(d), _p(tmp), len(a), 0, 64, _stream()))
    assert np.array_equal(d.cpu().numpy().view(np.uint64), np.sort(a))


@pytest.mark.parametrize("n", [10, 5000, 700_001])
def test_radix_sort_pairs_is_stable(n):
    L = _lib()
    rs = np.random.RandomState(n)
    k = rs.randint(0, 1000, size=n).astype(np.uint64)      # many equal keys -> stability is visible
    v = np.arange(n, dtype=np.uint32)
    dk = torch.from_numpy(k.view(np.int64)).cuda()
    dv = torch.from_numpy(v.view(np.int32)).cuda()
    tk, tv = torch.empty_like(dk), torch.empty_like(dv)
    L.check(L.lib().nbr_sort_pairs_u64_u32(_p(dk), _p(tk), _p(dv), _p(tv), n, 0, 16, _stream()))
    order = np.argsort(k, kind="stable")
    assert np.array_equal(dk.cpu().numpy().view(np.uint64), k[order])
    assert np.array_equal(dv.cpu().numpy().view(np.uint32), v[order])


def test_unique_sorted():
    L = _lib()
    a = np.sort(np.random.RandomState(0).randint(0, 50_000, size=400_000).astype(np.uint64))
    d = torch.from_numpy(a.view(np.int64)).cuda()
    out = torch.empty_like(d)
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    L.check(L.lib().nbr_unique_u64(_p(d), len(a), _p(out), _p(cnt), _stream()))
    ref = np.unique(a)
    assert int(cnt.item()) == len(ref)
    assert np.array_equal(out[:len(ref)].cpu().numpy().view(np.uint64), ref)


# ---------------------------------------------------------------------------------------------
# VoxelFilter mirror: the reference's own tests (nimrud/utils/tests/geometry_tests.py:17-279)
# ---------------------------------------------------------------------------------------------
BOUNDS = np.asarray([[0, 0, 0], [100, 100, 100]])


def test_voxel_init():
    from nimrud_b200 import geometry
    rs = np.random.RandomState(10)
    for dim in (2, 3):
        with pytest.raises(ValueError):
            geometry.VoxelFilter(rs.rand(1, dim) * 100, 0.5)
        pts = rs.rand(1000, dim) * 100
        vf = geometry.VoxelFilter(pts, 0.5)
        assert np.array_equal(vf.minimum_corner, pts.min(0) - 0.25)
        assert np.array_equal(vf.maximum_corner, pts.max(0) + 0.25)
        assert vf.edge_length == 0.5
    for dim in (1, 4):
        with pytest.raises(ValueError):
            geometry.VoxelFilter(rs.rand(1000, dim), 0.5)
    with pytest.raises(ValueError):
        geometry.VoxelFilter(rs.rand(10), 0.5)
    with pytest.raises(ValueError):
        geometry.VoxelFilter(rs.rand(10, 10, 10), 0.5)


def test_voxel_shift_and_masks():
    from nimrud_b200 import geometry
    for dim in (2, 3):
        vf = geometry.VoxelFilter(BOUNDS[:, :dim], 0.001)
        assert np.array_equal([17, 34][:dim - 1], vf.shifts)
        assert np.array_equal([17, 17, 17][:dim], vf.widths)
        with pytest.raises(ValueError):
            geometry.VoxelFilter(BOUNDS[:, :dim], 0.00001 if dim == 3 else 0.00000001)
        vf = geometry.VoxelFilter(BOUNDS[:, :dim], 1)
        assert np.array_equal([0b1111111, 0b11111110000000, 0b111111100000000000000][:dim], vf.masks)


def test_voxel_in_bounds():
    from nimrud_b200 import geometry
    for dim in (2, 3):
        vf = geometry.VoxelFilter(BOUNDS[:, :dim], 1)

        def ok(point):
            try:
                vf._check_in_bounds(point)
            except ValueError:
                return False
            return True

        assert ok(np.zeros((1, dim)) - 0.5)
        assert not ok(np.zeros((1, dim)) - 1.5)
        assert ok(np.zeros((1, dim)) + 0.5)
        assert ok(np.zeros((1, dim)) + 100.5)
        assert not ok(np.zeros((1, dim)) + 101.5)
        assert not ok(np.zeros((1, dim + 1)))
        assert ok(np.zeros(dim))
        assert not ok(np.zeros(dim + 1))


def test_voxel_address_and_transform():
    from nimrud_b200 import geometry
    vf = geometry.VoxelFilter(BOUNDS, 1)
    assert vf.coordinate_to_address(np.arange(3) + 10) == 198026
    assert np.allclose(np.arange(3) + 10, vf.address_to_coordinate(198026).flatten())
    vf2 = geometry.VoxelFilter(BOUNDS[:, :2], 1)
    known = (np.arange(3) + 10)[:2]
    assert np.allclose(known, vf2.address_to_coordinate(vf2.coordinate_to_address(known).flatten()))


def test_voxel_unique():
    from nimrud_b200 import geometry
    for dim in (2, 3):
        vf = geometry.VoxelFilter(BOUNDS[:, :dim], 1)
        pts = np.concatenate([np.zeros((1, dim)) + off for off in np.arange(0, 20, 2)])
        assert np.array_equal(pts, vf.unique_voxels(np.vstack((pts, pts))))


@pytest.mark.parametrize("name", ["small", "urban", "degenerate"])
def test_voxel_filter_matches_reference_bitwise(name):
    from nimrud_b200 import geometry
    g = load_golden(name)
    for dtype in (np.float32, np.float64):
        s = g["search"].astype(dtype)
        for i, e in enumerate(g["edges"]):
            vf = geometry.VoxelFilter(s, float(e))
            assert np.array_equal(vf.minimum_corner, g["s%d_min_corner" % i])
            assert np.array_equal(vf.maximum_corner, g["s%d_max_corner" % i])
            assert np.array_equal(vf.widths, g["s%d_widths" % i])
            assert np.array_equal(vf.shifts, g["s%d_shifts" % i])
            assert np.array_equal(vf.unique_addresses(s), g["s%d_addresses" % i])
            assert np.array_equal(vf.unique_voxels(s), g["s%d_centres" % i])


def test_voxel_filter_vs_oracle_1m(c_oracle):
    from nimrud_b200 import geometry, synth
    cloud = synth.urban_scene(1_000_000, seed=7).numpy()
    for e in (0.1, 0.4, 1.6):
        vf = geometry.VoxelFilter(cloud, e)
        minc, maxc, widths = c_oracle.grid_widths(cloud, e)
        assert np.array_equal(vf.minimum_corner, minc) and np.array_equal(vf.widths, widths)
        keys, centres = c_oracle.unique_voxels(cloud, minc, e, widths)
        assert np.array_equal(vf.unique_addresses(cloud), keys)
        assert np.array_equal(vf.unique_voxels(cloud), centres)
