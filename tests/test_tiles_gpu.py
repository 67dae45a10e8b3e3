"""
the multi-GPU tile + halo path, driven on ONE device: several tiles of one cloud live in one process, their halo
mailboxes are connected directly (nimrud_b200.distributed.process_tiles_local), every CUDA entry point of the
multi-rank run executes (nbr_tile_box_publish / nbr_tile_boxes_wait / nbr_halo_push / nbr_halo_wait /
nbr_order_cloud / nbr_multiscale_features_tile_mb, and the NCCL-transport trio nbr_halo_count / nbr_halo_fill /
nbr_multiscale_features_tile).  tile + halo must reproduce the unpartitioned call bit for bit, including the
queries within the halo width of a seam.  the selection rule itself is pinned to golden vectors produced by the
reference's nested_regions (nimrud/utils/geometry.py:203-253; its own test: utils/tests/geometry_tests.py:353-389).
"""
import ctypes
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden

pytestmark = pytest.mark.gpu

EDGES = (0.1, 0.2, 0.4, 0.8, 1.6)
RADII = (0.3, 0.6, 1.2, 2.4, 4.8)


def region_clouds():
    spec = importlib.util.spec_from_file_location("make_golden_regions", os.path.join(GOLDEN, "make_golden_regions.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.clouds()


def rows_sorted(t):
    a = t.detach().cpu().numpy().astype(np.float64)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def split_tiles(cloud, nx, ny):
    """nx x ny tiles by quantiles of x then y -> list of index tensors"""
    idx = []
    xq = torch.quantile(cloud[:, 0].double(), torch.linspace(0, 1, nx + 1, dtype=torch.float64, device=cloud.device))
    for i in range(nx):
        in_x = (cloud[:, 0] >= xq[i]) & ((cloud[:, 0] < xq[i + 1]) if i + 1 < nx else (cloud[:, 0] <= xq[i + 1]))
        sub = in_x.nonzero()[:, 0]
        yq = torch.quantile(cloud[sub, 1].double(), torch.linspace(0, 1, ny + 1, dtype=torch.float64, device=cloud.device))
        for j in range(ny):
            in_y = (cloud[sub, 1] >= yq[j]) & ((cloud[sub, 1] < yq[j + 1]) if j + 1 < ny else (cloud[sub, 1] <= yq[j + 1]))
            idx.append(sub[in_y])
    return idx


@pytest.fixture(scope="module")
def scene():
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(400_000, seed=31, device="cuda")
    whole = multiscale.process_single_core(cloud, cloud, EDGES, RADII, out_dtype=np.float32)
    return cloud, whole


@pytest.mark.parametrize("nx,ny", [(2, 1), (2, 2), (3, 2)])
def test_tiles_in_one_process_match_unpartitioned(scene, nx, ny):
    from nimrud_b200 import distributed as nd
    cloud, whole = scene
    idx = split_tiles(cloud, nx, ny)
    assert sum(i.numel() for i in idx) == cloud.shape[0]
    tiles = [cloud[i].contiguous() for i in idx]
    outs = nd.process_tiles_local(tiles, EDGES, RADII, out_dtype=np.float32)
    h = nd.halo_width(EDGES, RADII)
    seam_rows = 0
    for i, t, o in zip(idx, tiles, outs):
        assert torch.equal(o, whole[i]), "tile rows differ from the unpartitioned call"
        lo, hi = t.min(0).values, t.max(0).values
        seam_rows += int((((t[:, :2] - lo[:2]) < h) | ((hi[:2] - t[:, :2]) < h)).any(1).sum())
    assert seam_rows > 1000            # the comparison covered queries that need the halo


@pytest.mark.parametrize("mode", ["fused", "copy"])
def test_rows_gathered_by_the_feature_kernel(scene, mode):
    """the feature all-gather through peer stores: every tile's kernel writes its rows into the gather buffer of EVERY
    tile (here: four buffers on one device); each buffer must hold the unpartitioned rows in tile order.  "copy" is
    the path of scale sets the 7x7x7 kernel does not cover alone (rows into the own buffer, then peer copies)."""
    from nimrud_b200 import distributed as nd
    cloud, whole = scene
    idx = split_tiles(cloud, 2, 2)
    tiles = [cloud[i].contiguous() for i in idx]
    order = torch.cat(idx)
    if mode == "copy":
        edges, radii = (0.2, 0.2, 0.4), (0.6, 1.0, 1.2)        # 7x7x7 and 11x11x11 windows in one call
        from nimrud_b200 import multiscale
        ref = multiscale.process_single_core(cloud, cloud, edges, radii, out_dtype=np.float32)[order]
    else:
        edges, radii, ref = EDGES, RADII, whole[order]
    launches0 = _launches()
    outs = nd.process_tiles_local(tiles, edges, radii, out_dtype=np.float32, gather=True)
    assert len(outs) == 4 and _launches() > launches0
    for o in outs:
        assert o.shape == ref.shape and torch.equal(o, ref)
    # float64 rows, uneven tiles with an empty one
    tiles3 = [tiles[0], tiles[1][:0], torch.cat([tiles[1], tiles[2], tiles[3]])]
    order3 = torch.cat([idx[0], idx[1], idx[2], idx[3]])
    outs = nd.process_tiles_local(tiles3, edges, radii, out_dtype=np.float64, gather=True)
    if mode == "fused":
        from nimrud_b200 import multiscale
        ref64 = multiscale.process_single_core(cloud, cloud, edges, radii, out_dtype=np.float64)[order3]
        for o in outs:
            assert o.dtype == torch.float64 and torch.equal(o, ref64)


def _launches():
    from nimrud_b200 import _lib
    return int(_lib.lib().nbr_kernel_launches())


def test_empty_and_single_point_tiles(scene):
    from nimrud_b200 import distributed as nd
    cloud, whole = scene
    idx = split_tiles(cloud, 2, 1)
    lone = idx[1][:1]
    rest = idx[1][1:]
    tiles = [cloud[idx[0]].contiguous(), cloud[:0].contiguous(), cloud[lone].contiguous(), cloud[rest].contiguous()]
    outs = nd.process_tiles_local(tiles, EDGES, RADII, out_dtype=np.float32)
    assert outs[1].shape == (0, 4 * len(RADII))
    assert torch.equal(outs[0], whole[idx[0]])
    assert torch.equal(outs[2], whole[lone])
    assert torch.equal(outs[3], whole[rest])


def test_float64_tiles_and_float64_rows(scene):
    from nimrud_b200 import distributed as nd, multiscale
    cloud, _ = scene
    c64 = cloud[:150_000].double().contiguous()
    whole = multiscale.process_single_core(c64, c64, EDGES[:3], RADII[:3], out_dtype=np.float64)
    idx = split_tiles(c64, 2, 1)
    outs = nd.process_tiles_local([c64[i].contiguous() for i in idx], EDGES[:3], RADII[:3], out_dtype=np.float64)
    for i, o in zip(idx, outs):
        assert torch.equal(o, whole[i])


def test_mailbox_selection_matches_nested_regions_golden():
    """reference rule: search points inside [region_lo - buffer, region_hi + buffer], inclusive; a region without
    points selects nothing (geometry_tests.py:353-389).  tile 0 = the region's query points (its box IS the region:
    two of them sit on its corners), tile 1 = the search space, h = buffer radius."""
    from nimrud_b200 import distributed as nd
    g = load_golden("regions")
    query, search = region_clouds()
    q = torch.from_numpy(query[g["query_idx"]]).cuda().contiguous()
    s = torch.from_numpy(search).cuda().contiguous()
    assert np.array_equal(q.min(0).values.cpu().numpy(), g["lo"]) and np.array_equal(q.max(0).values.cpu().numpy(), g["hi"])
    boxes = nd.HaloMailbox.local_set(3, "cuda", torch.float64, 1 << 16)
    try:
        tiles = [q, s, s[:0]]                      # the third tile is a region that holds no point
        for mb, t in zip(boxes, tiles):
            mb.publish(t)
        for mb in boxes:
            mb.wait_boxes()
        for mb, t in zip(boxes, tiles):
            mb.push(t, float(g["buffer"]))
        from nimrud_b200 import _lib
        from nimrud_b200._util import stream_ptr
        for mb in boxes:
            _lib.check(_lib.lib().nbr_halo_wait(mb.handle, stream_ptr(mb.device)))
        got = boxes[0].received()
        assert got.shape[0] == g["search_idx"].size
        assert np.array_equal(rows_sorted(got), rows_sorted(torch.from_numpy(search[g["search_idx"]])))
        assert boxes[2].received().shape[0] == 0          # empty region: nothing selected
        assert g["empty_search_idx"].size == 0
    finally:
        for mb in boxes:
            mb.close()


def test_count_fill_selection_matches_nested_regions_golden():
    """the same rule through nbr_halo_count / nbr_halo_fill (the NCCL transport's selection kernels)."""
    from nimrud_b200 import distributed as nd
    g = load_golden("regions")
    _, search = region_clouds()
    s = torch.from_numpy(search).cuda().contiguous()
    b = float(g["buffer"])
    grown = [((g["lo"] - b).tolist(), (g["hi"] + b).tolist()), ([99.5] * 3, [110.5] * 3)]
    counts, state = nd._count_halos_cuda(s, grown)
    counts = [int(v) for v in counts.tolist()]
    assert counts == [g["search_idx"].size, 0]
    buf = nd._fill_halos_cuda(s, state, counts)
    assert np.array_equal(rows_sorted(buf), rows_sorted(torch.from_numpy(search[g["search_idx"]])))


def test_nccl_transport_entry_points_on_one_device(scene):
    """nbr_halo_count -> nbr_halo_fill -> (device copy = the exchange) -> nbr_order_cloud -> nbr_multiscale_features_tile
    for two tiles, against the unpartitioned call."""
    from nimrud_b200 import _lib, distributed as nd
    from nimrud_b200._util import ptr, stream_ptr
    cloud, whole = scene
    idx = split_tiles(cloud, 2, 1)
    tiles = [cloud[i].contiguous() for i in idx]
    h = nd.halo_width(EDGES, RADII)
    box = [(t.min(0).values.double().cpu().numpy(), t.max(0).values.double().cpu().numpy()) for t in tiles]
    g_lo = np.minimum(box[0][0], box[1][0])
    g_hi = np.maximum(box[0][1], box[1][1])
    f64p = ctypes.POINTER(ctypes.c_double)
    lib = _lib.lib()
    for me, other in ((0, 1), (1, 0)):
        grown = [((box[me][0] - h).tolist(), (box[me][1] + h).tolist())]
        counts, state = nd._count_halos_cuda(tiles[other], grown)
        halo = nd._fill_halos_cuda(tiles[other], state, [int(v) for v in counts.tolist()])
        glob = np.ascontiguousarray(np.concatenate([g_lo, g_hi]))
        local = np.ascontiguousarray(np.concatenate([np.maximum(box[me][0] - h, g_lo), np.minimum(box[me][1] + h, g_hi)]))
        mine = np.ascontiguousarray(np.concatenate(box[me]))
        origin = np.zeros(3)
        _lib.check(lib.nbr_brick_origin(glob.ctypes.data_as(f64p), local.ctypes.data_as(f64p), min(EDGES), origin.ctypes.data_as(f64p)))
        n = tiles[me].shape[0]
        perm = torch.empty(n, dtype=torch.int32, device="cuda")
        ordered = torch.empty_like(tiles[me])
        _lib.check(lib.nbr_order_cloud(ptr(tiles[me]), _lib.F32, n, mine.ctypes.data_as(f64p), origin.ctypes.data_as(f64p),
                                       min(EDGES), ptr(perm), ptr(ordered), stream_ptr()))
        out = torch.zeros((n, 4 * len(RADII)), dtype=torch.float32, device="cuda")
        e_arr, e_p = _lib.f64_array(EDGES)
        r_arr, r_p = _lib.f64_array(RADII)
        _lib.check(lib.nbr_multiscale_features_tile(ptr(ordered), ptr(perm), _lib.F32, n, ptr(halo), halo.shape[0],
                                                    local.ctypes.data_as(f64p), glob.ctypes.data_as(f64p), e_p, r_p, len(RADII),
                                                    ptr(out), _lib.F32, 0, None, stream_ptr()))
        assert torch.equal(out, whole[idx[me]])


def test_mailbox_overflow_is_reported(scene):
    from nimrud_b200 import distributed as nd
    cloud, _ = scene
    idx = split_tiles(cloud, 2, 1)
    tiles = [cloud[i].contiguous() for i in idx]
    with pytest.raises(RuntimeError, match="dropped"):
        nd.process_tiles_local(tiles, EDGES, RADII, capacity_rows=100)


def test_tile_step_with_host_buffers(scene):
    """nbr_tile_step_host on a world of one tile: rows == the device-resident call (float32 on the wire)."""
    from nimrud_b200 import distributed as nd
    cloud, whole = scene
    mb = nd.HaloMailbox(0, 1, "cuda", torch.float32, 1024)
    try:
        host = cloud.cpu().numpy()
        rows = nd.process_tile_host(host, EDGES, RADII, out_dtype=np.float64, device="cuda:0", mailbox=mb)
        assert rows.dtype == np.float64 and np.array_equal(rows, whole.cpu().numpy().astype(np.float64))
        rows32 = nd.process_tile_host(torch.from_numpy(host).pin_memory(), EDGES, RADII, out_dtype=np.float32, device="cuda:0", mailbox=mb)
        assert np.array_equal(rows32, whole.cpu().numpy())
    finally:
        mb.close()
