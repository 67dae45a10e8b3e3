"""
world_size-2 gloo test of the multi-GPU host logic (box reduction, halo selection, all-to-all-v, gather):
tile + halo with globally anchored lattices must reproduce the unpartitioned result bit for bit.
the compute function injected here is the oracle (the CUDA path cannot run in this container).
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

EDGES = (0.2, 0.5)
RADII = (0.6, 1.5)


def anchored_oracle(query, search, edges, radii, bbox, out_dtype, out):
    """oracle features with the voxel lattice anchored on `bbox` instead of on search's own box."""
    from oracle import nimrud_oracle as O
    q = query.numpy().astype(np.float64)
    s = search.numpy().astype(np.float64)
    blocks = []
    for e, r in zip(edges, radii):
        params = O.GridParams(bbox[0] - e / 2, bbox[1] + e / 2, e)
        _, centres = O.unique_voxels(params, s)
        off, idx = O.radius_sets(q, centres, r)
        blocks.append(O.rows_from_sets(q, centres, off, idx))
    return torch.from_numpy(np.concatenate(blocks, axis=1))


def make_cloud():
    rs = np.random.RandomState(7)
    pts = rs.rand(1600, 3) * [8.0, 4.0, 1.5]
    return pts.astype(np.float32)


def worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nimrud_b200 import distributed as nd
    cloud = make_cloud()
    left = cloud[:, 0] < 4.0
    mine = torch.from_numpy(cloud[left] if rank == 0 else cloud[~left])
    halo, (g_lo, g_hi), _ = nd.exchange_halo(mine, EDGES, RADII)
    # the halo is exactly the foreign points within h of my box
    other = torch.from_numpy(cloud[~left] if rank == 0 else cloud[left])
    lo, hi = nd.tile_box(mine)
    expect = other[nd.select_halo(other, lo, hi, nd.halo_width(EDGES, RADII))]
    assert halo.shape == expect.shape and torch.equal(halo, expect)
    assert torch.equal(g_lo, torch.from_numpy(cloud.astype(np.float64).min(0)))
    assert torch.equal(g_hi, torch.from_numpy(cloud.astype(np.float64).max(0)))
    feats = nd.process_tile(mine, EDGES, RADII, gather=True, compute=anchored_oracle)
    try:
        nd.process_tile(mine, EDGES, RADII, gather="peer", compute=anchored_oracle)      # peer stores need the CUDA path
        raise AssertionError("gather='peer' on CPU tensors must be refused")
    except ValueError:
        pass
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), feats.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_tile_plus_halo_equals_unpartitioned(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    from oracle import nimrud_oracle as O
    cloud = make_cloud().astype(np.float64)
    left = cloud[:, 0] < 4.0
    order = np.concatenate([np.nonzero(left)[0], np.nonzero(~left)[0]])
    ref = O.process(cloud, cloud, EDGES, RADII)[order]
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)          # bit for bit: same voxels, same neighbor sets


def test_gather_mode_rule():
    """which transport the final feature all-gather takes (no collective needed to decide)"""
    from nimrud_b200.distributed import gather_mode
    assert gather_mode(False, 8, True, env="") is None
    assert gather_mode(True, 2, True, env="") == "peer" and gather_mode(True, 4, True, env="") == "peer"
    assert gather_mode(True, 8, True, env="") == "nccl"                 # the receivers' placement pass outweighs the overlap
    assert gather_mode(True, 8, True, env="peer") == "peer" and gather_mode(True, 2, True, env="nccl") == "nccl"
    assert gather_mode(True, 2, True, env="copy") == "peer"             # the push-kernel form of the peer path
    assert gather_mode("nccl", 2, True, env="peer") == "nccl" and gather_mode("peer", 8, True, env="nccl") == "peer"
    assert gather_mode(True, 2, False, env="peer") == "nccl"            # CPU tensors / custom compute: collective gather
    with pytest.raises(ValueError):
        gather_mode("peer", 2, False, env="")


def test_halo_width_rule():
    from nimrud_b200 import distributed as nd
    assert nd.halo_width((0.1, 1.6), (0.3, 4.8)) >= 4.8 + 0.8


def test_select_halo_matches_nested_regions_golden():
    """the host selection rule against golden vectors from the reference's nested_regions
    (nimrud/utils/geometry.py:203-253): inclusive box +- buffer radius; a region without points selects nothing."""
    import importlib.util
    from conftest import GOLDEN, load_golden
    from nimrud_b200 import distributed as nd
    spec = importlib.util.spec_from_file_location("make_golden_regions", os.path.join(GOLDEN, "make_golden_regions.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    query, search = mod.clouds()
    g = load_golden("regions")
    lo, hi = torch.from_numpy(g["lo"]), torch.from_numpy(g["hi"])
    for dtype in (torch.float64, torch.float32):      # the float32 path rounds the bounds outward: same set here
        idx = nd.select_halo(torch.from_numpy(search).to(dtype), lo, hi, float(g["buffer"]))
        assert np.array_equal(idx.numpy(), g["search_idx"])
    idx = nd.select_halo(torch.from_numpy(query), lo, hi, 0.0)
    assert np.array_equal(idx.numpy(), g["query_idx"])
    far = torch.full((3,), 100.0, dtype=torch.float64)
    assert nd.select_halo(torch.from_numpy(search), far, far + 10, float(g["buffer"])).numel() == 0
