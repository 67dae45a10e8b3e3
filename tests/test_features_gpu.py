"""
GPU parity of the fused radius-feature path against golden vectors produced by the reference, the
oracle on seeded clouds, and size-independent properties.  all calls go through the C ABI.
"""
import numpy as np
import pytest
import torch

from conftest import assert_features_close, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["small", "urban", "degenerate"])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_process_single_core_matches_reference(name, dtype):
    from nimrud_b200 import multiscale
    g = load_golden(name)
    q, s = g["query"].astype(dtype), g["search"].astype(dtype)
    out = multiscale.process_single_core(q, s, tuple(g["edges"]), tuple(g["radii"]))
    assert isinstance(out, np.ndarray) and out.dtype == np.float64
    assert out.shape == g["features"].shape
    assert_features_close(out, g["features"], g["radii"])
    # float32 output option
    out32 = multiscale.process_single_core(q, s, tuple(g["edges"]), tuple(g["radii"]), out_dtype=np.float32)
    assert out32.dtype == np.float32
    assert np.array_equal(out32[:, 0::4], g["features"][:, 0::4])
    assert np.allclose(out32, g["features"], rtol=1e-4, atol=1e-5)


def test_one_scale_and_device_tensors():
    from nimrud_b200 import multiscale
    g = load_golden("small")
    q = torch.from_numpy(g["query"]).cuda()
    out = multiscale.one_scale_single_core(q, q, float(g["edges"][1]), float(g["radii"][1]))
    assert out.is_cuda and out.shape == (len(g["query"]), 4) and out.dtype == torch.float64
    assert_features_close(out.cpu().numpy(), g["features"][:, 4:8], g["radii"][1:2])


def test_config1_subset_matches_reference():
    from nimrud_b200 import multiscale, synth
    g = load_golden("config1")
    cloud = synth.uniform_box()
    out, counts = multiscale.process_single_core(cloud[g["pick"]], cloud, tuple(g["edges"]), tuple(g["radii"]),
                                                 return_voxel_counts=True)
    assert np.array_equal(counts, g["n_voxels"])        # 94187 / 65061 / 15449
    assert_features_close(out, g["features"], g["radii"])


@pytest.mark.parametrize("name", ["small", "urban", "degenerate"])
def test_radius_sets_bit_exact(name):
    from nimrud_b200 import multiscale
    g = load_golden(name)
    q, s = g["query"], g["search"]
    for i, (e, r) in enumerate(zip(g["edges"], g["radii"])):
        index = multiscale.LatticeIndex(s, float(e), indexed=True)
        assert index.n_voxels == len(g["s%d_addresses" % i])
        addr, centres = index.addresses_and_centres()
        assert np.array_equal(addr.cpu().numpy(), g["s%d_addresses" % i])
        assert np.array_equal(centres.cpu().numpy(), g["s%d_centres" % i])
        off, idx = index.radius_sets(q, float(r))
        assert np.array_equal(off.cpu().numpy(), g["s%d_offsets" % i])
        assert np.array_equal(idx.cpu().numpy(), g["s%d_indices" % i])
        index.close()


def test_exact_and_row_kernels_agree_with_each_other_and_the_oracle(c_oracle):
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(300_000, seed=3).numpy()
    q = cloud[::7]
    for e, radii in ((0.2, (0.4, 0.6, 0.8, 1.0, 1.2)), (0.4, (1.2,)), (1.6, (4.8,))):
        index = multiscale.LatticeIndex(cloud, e)
        a = index.radius_features(q, radii, algorithm=1).cpu().numpy()
        b = index.radius_features(q, radii, algorithm=0).cpu().numpy()
        ref = c_oracle.process(q, cloud, [e] * len(radii), radii, threads=8)
        assert_features_close(a, ref, radii)
        assert_features_close(b, ref, radii)
        index.close()


def test_boundary_inclusive_and_lattice_aligned_queries(c_oracle):
    # queries sitting exactly on voxel centres with r an exact multiple of e: distance == r ties
    from nimrud_b200 import multiscale
    rs = np.random.RandomState(0)
    cells = rs.randint(0, 12, size=(4000, 3))
    search = (cells * 0.5 + 0.25).astype(np.float32)
    q = search[:500].copy()
    for r in (0.5, 1.0, 1.5):
        ref = c_oracle.process(q, search, [0.5], [r])
        out = multiscale.process_single_core(q, search, [0.5], [r])
        assert_features_close(out, ref, [r])


def test_error_behaviour_matches_reference():
    from nimrud_b200 import multiscale
    pts = np.random.RandomState(0).rand(50, 3)
    with pytest.raises(AssertionError):
        multiscale.process_single_core(pts, pts, [0.1, 0.2], [0.5])
    with pytest.raises(ValueError):
        multiscale.process_single_core(pts, pts[:1], [0.1], [0.5])
    with pytest.raises(ValueError):                      # > 64 address bits (utils/geometry.py:59-60)
        multiscale.process_single_core(pts * 1e6, pts * 1e6, [1e-9], [0.5])
    out = multiscale.process_single_core(pts[:0], pts, [0.1], [0.5])
    assert out.shape == (0, 4)


def test_query_sharding_invariance():
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(200_000, seed=5).numpy()
    e, r = (0.2, 0.8), (0.6, 2.4)
    whole = multiscale.process_single_core(cloud, cloud, e, r)
    parts = np.concatenate([multiscale.process_single_core(cloud[a:a + 50_000], cloud, e, r)
                            for a in range(0, 200_000, 50_000)])
    assert np.array_equal(whole, parts)


def test_full_size_properties_10m():
    # BASELINE config 2 at full size: size-independent properties instead of the oracle
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(10_000_000, seed=20, device="cuda")
    edges = (0.1, 0.2, 0.4, 0.8, 1.6)
    radii = (0.3, 0.6, 1.2, 2.4, 4.8)
    out = multiscale.process_single_core(cloud, cloud, edges, radii, out_dtype=np.float32)
    assert out.shape == (10_000_000, 20)
    assert torch.isfinite(out).all()
    pop = out[:, 0::4]
    assert (pop >= 1).all()                       # query == search: own voxel is always a neighbor
    assert (pop == pop.round()).all()
    l1, l2 = out[:, 2::4], out[:, 3::4]
    assert (l1 >= l2 - 1e-6).all() and (l2 >= -1e-6).all() and (l1 + l2 <= 1 + 1e-5).all()
    assert (l1 >= 1.0 / 3 - 1e-5).logical_or(pop < 2).all()
    cen = out[:, 1::4]
    assert (cen <= torch.tensor(radii, device="cuda") * (1 + 1e-5)).all()
    # a random 20k subset against the same call on the subset (sharding invariance at full size)
    pick = torch.randperm(10_000_000, device="cuda")[:20_000]
    sub = multiscale.process_single_core(cloud[pick].contiguous(), cloud, edges, radii, out_dtype=np.float32)
    assert torch.equal(sub, out[pick])


def test_subset_of_10m_against_oracle(c_oracle):
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(2_000_000, seed=20, device="cuda")
    edges = (0.1, 0.2, 0.4, 0.8, 1.6)
    radii = (0.3, 0.6, 1.2, 2.4, 4.8)
    pick = torch.randperm(2_000_000, device="cuda")[:30_000]
    out = multiscale.process_single_core(cloud[pick].contiguous(), cloud, edges, radii).cpu().numpy()
    host = cloud.cpu().numpy()
    ref = c_oracle.process(host[pick.cpu().numpy()], host, edges, radii, threads=8)
    assert_features_close(out, ref, radii)


def test_extended_descriptors_against_oracle():
    from nimrud_b200 import multiscale
    from oracle import nimrud_oracle as O
    g = load_golden("small")
    q = g["query"][:300]
    s = g["search"]
    e, r = float(g["edges"][1]), float(g["radii"][1])
    out = multiscale.process_single_core(q, s, [e], [r], descriptors="extended")
    assert out.shape == (300, 26)
    assert_features_close(out[:, :4], g["features"][:300, 4:8], [r])
    p = O.grid_params(s.astype(np.float64), e)
    _, centres = O.unique_voxels(p, s.astype(np.float64))
    off, idx = O.radius_sets(q.astype(np.float64), centres, r)
    ext = O.rows_from_sets(q.astype(np.float64), centres, off, idx, row_fn=O.extended_row)
    assert np.allclose(out[:, 4:12], ext[:, :8], rtol=1e-4, atol=1e-7)
    assert np.allclose(out[:, 15], ext[:, 11], rtol=1e-6)
    assert np.allclose(out[:, 16:22], ext[:, 12:18], rtol=1e-6, atol=1e-12 + 1e-9 * np.abs(ext[:, 11:12]))   # covariance
    # normals: compare where the smallest eigenvalue is well separated
    sep = (ext[:, 1] > 0.05)            # planarity = (e2-e3)/e1
    dots = np.abs((out[sep, 12:15] * ext[sep, 8:11]).sum(1))
    assert (dots > 1 - 1e-6).all()
    # x, y of the eigenvectors of the largest and the middle eigenvalue: where all three eigenvalues are well
    # separated and the sign convention (x > 0) is not decided by rounding
    clear = (ext[:, 0] > 0.05) & (ext[:, 1] > 0.05)
    lead_ok = clear & (np.abs(ext[:, 18]) > 1e-3)
    second_ok = clear & (np.abs(ext[:, 20]) > 1e-3)
    assert lead_ok.sum() > 50 and second_ok.sum() > 50
    assert np.allclose(out[lead_ok, 22:24], ext[lead_ok, 18:20], atol=1e-4)
    assert np.allclose(out[second_ok, 24:26], ext[second_ok, 20:22], atol=1e-4)


def test_launch_counter_moves():
    from nimrud_b200 import _lib, multiscale
    before = _lib.lib().nbr_kernel_launches()
    pts = np.random.RandomState(0).rand(1000, 3).astype(np.float32)
    multiscale.process_single_core(pts, pts, [0.1], [0.3])
    assert _lib.lib().nbr_kernel_launches() > before


def test_raw_ctypes_stub_from_integration_md():
    # the binding INTEGRATION.md shows a maintainer of the reference: plain ctypes, host buffers, float64
    import ctypes
    from nimrud_b200 import _lib
    g = load_golden("small")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    lib.nbr_last_error.restype = ctypes.c_char_p
    f64p = ctypes.POINTER(ctypes.c_double)
    q = np.ascontiguousarray(g["query"], dtype=np.float64)
    e = np.ascontiguousarray(g["edges"], dtype=np.float64)
    r = np.ascontiguousarray(g["radii"], dtype=np.float64)
    out = np.zeros((len(q), 4 * len(r)))
    rc = lib.nbr_multiscale_features_host(
        q.ctypes.data_as(ctypes.c_void_p), 1, ctypes.c_int64(len(q)), q.ctypes.data_as(ctypes.c_void_p), 1,
        ctypes.c_int64(len(q)), e.ctypes.data_as(f64p), r.ctypes.data_as(f64p), ctypes.c_int32(len(r)),
        out.ctypes.data_as(ctypes.c_void_p), 1, ctypes.c_int32(0), None)
    assert rc == 0, lib.nbr_last_error()
    assert_features_close(out, g["features"], g["radii"])
    # separate query / search arrays and several batches through the pipelined host path
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(700_000, seed=9).numpy()
    a = multiscale.process_single_core(cloud[:300_000].copy(), cloud, (0.2, 0.4), (0.6, 1.2))
    dev_q, dev_s = torch.from_numpy(cloud[:300_000]).cuda(), torch.from_numpy(cloud).cuda()
    b32 = multiscale.process_single_core(dev_q, dev_s, (0.2, 0.4), (0.6, 1.2), out_dtype=np.float32).cpu().numpy()
    # host buffers: the rows cross PCIe as float32 and are widened on the host (include/nimrud_b200.h)
    assert a.dtype == np.float64 and np.array_equal(a, b32.astype(np.float64))
    # NBR_HOST_WIRE=f64 keeps float64 on the wire: identical to the device-resident float64 rows
    import os
    os.environ["NBR_HOST_WIRE"] = "f64"
    try:
        a64 = multiscale.process_single_core(cloud[:300_000].copy(), cloud, (0.2, 0.4), (0.6, 1.2))
    finally:
        os.environ.pop("NBR_HOST_WIRE")
    b = multiscale.process_single_core(dev_q, dev_s, (0.2, 0.4), (0.6, 1.2)).cpu().numpy()
    assert np.array_equal(a64, b)
    # float32 rows into a pageable float32 result
    a32 = multiscale.process_single_core(cloud[:300_000].copy(), cloud, (0.2, 0.4), (0.6, 1.2), out_dtype=np.float32)
    assert np.array_equal(a32, b32)


def test_mixed_window_widths_in_one_call(c_oracle):
    # r/e = 2, 3 (7x7x7 windows: shell-table kernel, equal edges share one staged window), r/e = 5
    # (interval kernel) and r/e = 12 (per-candidate kernel) in ONE call: every column block comes from a
    # different kernel, none may touch the others' columns.  device tensors and host arrays.
    import torch
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(150_000, seed=5).numpy()
    q = cloud[::5].copy()
    edges = (0.2, 0.2, 0.2, 0.1)
    radii = (0.4, 0.6, 1.0, 1.2)
    ref = c_oracle.process(q, cloud, edges, radii, threads=8)
    host = multiscale.process_single_core(q, cloud, edges, radii)
    assert_features_close(host, ref, radii)
    dev = multiscale.process_single_core(torch.from_numpy(q).cuda(), torch.from_numpy(cloud).cuda(), edges, radii,
                                         out_dtype=np.float32)
    assert_features_close(dev.cpu().numpy(), ref, radii)


def test_shell_table_exactness_on_a_dense_lattice(c_oracle):
    # every cell of a block occupied: all 343 cells of every window are candidates, queries at arbitrary
    # fractional positions and exactly on centres / faces / corners, radii that put many centres exactly on
    # the sphere.  populations must match the float64 oracle exactly.
    from nimrud_b200 import multiscale
    rs = np.random.RandomState(11)
    g = np.stack(np.meshgrid(np.arange(14), np.arange(14), np.arange(14), indexing="ij"), -1).reshape(-1, 3)
    search = (g * 0.25 + 0.125).astype(np.float32)
    inner = search[(g.min(1) >= 4) & (g.max(1) <= 9)]
    q = np.concatenate([inner[:200], inner[:200] + 0.125, inner[:200] + np.float32(0.0625),
                        (rs.rand(3000, 3) * 1.25 + 1.0).astype(np.float32)]).astype(np.float32)
    for r in (0.25, 0.5, 0.75, 0.559017, 0.8291562):          # e, 2e, 3e, sqrt(5) e, sqrt(11) e
        ref = c_oracle.process(q, search, [0.25], [r])
        out = multiscale.process_single_core(q, search, [0.25], [r])
        assert np.array_equal(out[:, 0], ref[:, 0]), r
        assert_features_close(out, ref, [r])


def test_far_and_outside_queries(c_oracle):
    # queries outside the search cloud's box (partly or completely): windows hang over the lattice edge
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(60_000, seed=6).numpy()
    lo, hi = cloud.min(0), cloud.max(0)
    rs = np.random.RandomState(4)
    q = np.concatenate([cloud[:500], lo - rs.rand(300, 3).astype(np.float32), hi + rs.rand(300, 3).astype(np.float32),
                        lo - 1000.0 + rs.rand(50, 3).astype(np.float32), [[1e6, -1e6, 3.0]]]).astype(np.float32)
    edges, radii = (0.2, 0.4), (0.6, 1.2)
    ref = c_oracle.process(q, cloud, edges, radii)
    out = multiscale.process_single_core(q, cloud, edges, radii)
    assert_features_close(out, ref, radii)


def test_queries_as_prefix_of_the_search_buffer(c_oracle):
    # the multi-GPU tile + halo hand-off: the query tensor is a view of the first rows of the search tensor.
    # the CUDA path then orders the whole buffer once and keeps only the prefix as queries.
    import torch
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(90_000, seed=8)
    dev = cloud.cuda()
    edges, radii = (0.2, 0.4, 0.2), (0.6, 1.2, 1.0)           # shell-table kernel + interval kernel (r/e = 5)
    for nq in (1, 31, 60_000):
        out = multiscale.process_single_core(dev[:nq], dev, edges, radii, out_dtype=np.float32)
        assert out.shape == (nq, 12)
        ref = c_oracle.process(cloud.numpy()[:nq], cloud.numpy(), edges, radii, threads=8)
        assert_features_close(out.cpu().numpy(), ref, radii)
    # and the same rows as the full-cloud call
    full = multiscale.process_single_core(dev, dev, edges, radii, out_dtype=np.float32)
    part = multiscale.process_single_core(dev[:60_000], dev, edges, radii, out_dtype=np.float32)
    assert torch.equal(full[:60_000], part)


def test_many_distinct_radius_ratios_recycle_the_table_caches(c_oracle):
    # every distinct r/e builds a shell table; beyond 128 cached tables the caches are dropped and rebuilt
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(8_000, seed=12).numpy()
    q = cloud[:1500]
    checked = 0
    for i in range(150):
        r3, r5 = 0.40 + 0.002 * i, 0.80 + 0.002 * i           # r/e in [2, 3.5) and [4, 5.5) at e = 0.2
        out = multiscale.process_single_core(q, cloud, [0.2, 0.2], [r3, r5])
        if i % 37 == 0 or i >= 147:
            ref = c_oracle.process(q, cloud, [0.2, 0.2], [r3, r5])
            assert_features_close(out, ref, [r3, r5])
            checked += 1
    assert checked >= 7


def test_table_kernels_against_the_per_candidate_kernel_at_scale():
    # two independent code paths on 2M queries of the 10M-point scene: the shell-table kernels decide most
    # cells by table lookup, the per-candidate kernel evaluates the reference's float64 expression for every
    # occupied cell of the window.  populations (neighbor set sizes) must be identical; the eigen ratios come from two
    # different solvers (cubic + deflated quadratic on exact integer minors vs eigenvector deflation) and must agree
    # 100x tighter than the parity bar: a double root of the quadratic costs sqrt(1e-16) = 1e-8 absolute.
    import torch
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(10_000_000, seed=20, device="cuda")
    q = cloud[::5].contiguous()
    for e, radii in ((0.1, (0.3,)), (0.4, (1.2, 2.0)), (1.6, (4.8, 8.0))):
        index = multiscale.LatticeIndex(cloud, e)
        fast = index.radius_features(q, radii, out_dtype=np.float64, algorithm=0)
        slow = index.radius_features(q, radii, out_dtype=np.float64, algorithm=1)
        index.close()
        assert torch.equal(fast[:, 0::4], slow[:, 0::4]), "populations differ at e = %g" % e
        d = (fast - slow).abs()
        for k, r in enumerate(radii):
            assert d[:, 4 * k + 1].max().item() <= 1e-5 * r          # centroid: float32 vs float64 finish
            lim = 1e-6 * slow[:, 4 * k + 2:4 * k + 4].abs() + 1e-9
            assert bool((d[:, 4 * k + 2:4 * k + 4] <= lim).all())      # same integer moments, two eigen-solvers


def test_many_scales_in_one_call(c_oracle):
    # 18 scales: more entries than one launch of the table kernels holds and rows wider than the row buffer
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(30_000, seed=15).numpy()
    q = cloud[:2500]
    edges = [0.1] * 6 + [0.2] * 6 + [0.4] * 6
    radii = [e * f for e, f in zip(edges, [1.0, 1.5, 2.0, 2.5, 3.0, 3.4] * 3)]
    ref = c_oracle.process(q, cloud, edges, radii, threads=8)
    for dt in (np.float32, np.float64):
        out = multiscale.process_single_core(q, cloud, edges, radii, out_dtype=dt)
        assert out.shape == (2500, 72)
        assert_features_close(out, ref, radii)


def test_large_extent_sparse_cloud_is_tiled():
    """a 10 km x 10 km x 200 m cloud at e = 0.1 fits the reference's 64-bit voxel address (17 + 17 + 11 bits,
    utils/geometry.py:55-60) but not a dense brick directory: the shim processes it tile by tile with the lattices
    anchored on the whole cloud's box.  checked against the oracle; numpy and CUDA-tensor inputs."""
    from nimrud_b200 import multiscale
    from oracle import nimrud_oracle as O
    rs = np.random.RandomState(4)
    a = rs.rand(3000, 3) * [25.0, 25.0, 3.0]
    b = rs.rand(3000, 3) * [25.0, 25.0, 3.0] + [9990.0, 9970.0, 197.0]
    lone = np.array([[5000.0, 5000.0, 100.0]])
    cloud = np.concatenate([a, b, lone]).astype(np.float32)
    edges, radii = (0.1, 0.4), (0.3, 1.2)
    with pytest.raises(NotImplementedError):
        multiscale.LatticeIndex(cloud, 0.1)                       # the dense directory itself still refuses
    got = multiscale.process_single_core(cloud, cloud, edges, radii)
    ref = O.process(cloud.astype(np.float64), cloud.astype(np.float64), edges, radii)
    assert_features_close(got, ref, radii)
    assert got[-1, 0] == 1.0 and got[-1, 4] == 1.0                # the lone point sees its own voxel
    dev = torch.from_numpy(cloud).cuda()
    got_dev = multiscale.process_single_core(dev, dev, edges, radii, out_dtype=np.float32)
    assert np.array_equal(got_dev.cpu().numpy()[:, 0::4], ref[:, 0::4].astype(np.float32))
    # separate query cloud, some queries outside the search cloud's box
    q = np.concatenate([cloud[::7], [[-0.2, 3.0, 1.0], [12000.0, 0.0, 0.0]]]).astype(np.float32)
    got_q = multiscale.process_single_core(q, cloud, edges, radii)
    ref_q = O.process(q.astype(np.float64), cloud.astype(np.float64), edges, radii)
    assert_features_close(got_q, ref_q, radii)
