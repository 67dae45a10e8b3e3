"""2-GPU NCCL run of the tile + halo path against the single-GPU result (skipped with < 2 GPUs)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

pytestmark = pytest.mark.gpu

EDGES = (0.1, 0.2, 0.4, 0.8, 1.6)
RADII = (0.3, 0.6, 1.2, 2.4, 4.8)
N = 400_000


def worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from nimrud_b200 import distributed as nd, synth
    cloud = synth.urban_scene(N, seed=31, device="cpu")
    split = cloud[:, 0].median()
    mine = cloud[cloud[:, 0] < split] if rank == 0 else cloud[cloud[:, 0] >= split]
    feats = nd.process_tile(mine.cuda().contiguous(), EDGES, RADII, gather=True, out_dtype=np.float32)
    # the same through the NCCL transport (all-gather of the boxes, all-to-all-v of counts and points)
    os.environ["NBR_HALO"] = "nccl"
    feats_nccl = nd.process_tile(mine.cuda().contiguous(), EDGES, RADII, gather=True, out_dtype=np.float32)
    os.environ.pop("NBR_HALO")
    assert torch.equal(feats, feats_nccl)
    # the all-gather through peer stores from inside the feature kernel (nbr_tile_step_gather), twice (the second step
    # reuses the buffers), and with a scale set the 7x7x7 kernel does not cover alone (peer copies of the finished share)
    for _ in range(2):
        peer = nd.process_tile(mine.cuda().contiguous(), EDGES, RADII, gather="peer", out_dtype=np.float32)
        assert torch.equal(peer, feats)
    mixed_e, mixed_r = (0.2, 0.2, 0.4), (0.6, 1.0, 1.2)
    a = nd.process_tile(mine.cuda().contiguous(), mixed_e, mixed_r, gather="nccl", out_dtype=np.float32)
    b = nd.process_tile(mine.cuda().contiguous(), mixed_e, mixed_r, gather="peer", out_dtype=np.float32)
    assert torch.equal(a, b)
    b64 = nd.process_tile(mine.cuda().contiguous(), EDGES, RADII, gather="peer", out_dtype=np.float64)
    a64 = nd.process_tile(mine.cuda().contiguous(), EDGES, RADII, gather="nccl", out_dtype=np.float64)
    assert b64.dtype == torch.float64 and torch.equal(a64, b64)
    keep = torch.empty((N + 7, 20), dtype=torch.float32, device="cuda")
    got = nd.process_tile(mine.cuda().contiguous(), EDGES, RADII, gather=True, out_all=keep)
    assert got.data_ptr() == keep.data_ptr() and torch.equal(got, feats)
    # host buffers (nbr_tile_step_host): float32 on the wire, widened on the host
    local = nd.process_tile(mine.cuda().contiguous(), EDGES, RADII, out_dtype=np.float32)
    host_rows = nd.process_tile_host(mine.numpy(), EDGES, RADII, out_dtype=np.float64)
    assert host_rows.dtype == np.float64 and np.array_equal(host_rows, local.cpu().numpy().astype(np.float64))
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), feats.cpu().numpy())
    dist.barrier()
    nd.release_mailboxes()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_tiles_match_single_gpu(tmp_path):
    from nimrud_b200 import multiscale, synth
    port = 29600 + (os.getpid() % 2000)
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    cloud = synth.urban_scene(N, seed=31, device="cpu")
    split = cloud[:, 0].median()
    left = cloud[:, 0] < split
    order = torch.cat([left.nonzero()[:, 0], (~left).nonzero()[:, 0]])
    whole = multiscale.process_single_core(cloud.cuda(), cloud.cuda(), EDGES, RADII, out_dtype=np.float32)
    ref = whole[order.cuda()].cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)     # same voxels, same integer moments, same arithmetic -> identical


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_devices_in_one_process():
    """kernel attributes (dynamic shared memory limits) and table caches are per device: one process may
    call the path on cuda:0 and then on cuda:1 (radius kernels of every window size, kNN, radix sort)."""
    from nimrud_b200 import multiscale, synth
    cloud = synth.urban_scene(60_000, seed=5, device="cpu")
    edges, radii = (0.2, 0.2, 0.2), (0.6, 1.0, 1.4)          # 7x7x7, 11x11x11 and interval kernels
    outs, knns = [], []
    for d in (0, 1):
        c = cloud.to("cuda:%d" % d)
        outs.append(multiscale.process_single_core(c, c, edges, radii, out_dtype=np.float32).cpu())
        knns.append(multiscale.knn_features(c, c, 0.2, (5, 10), out_dtype=np.float32).cpu())
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(knns[0], knns[1])
    # the voxel filter (bbox, addresses, radix sort + unique, centres) on the second device
    from nimrud_b200.geometry import VoxelFilter
    cen = [VoxelFilter(cloud.to("cuda:%d" % d), 0.4).unique_voxels(cloud.to("cuda:%d" % d)).cpu() for d in (0, 1)]
    assert torch.equal(cen[0], cen[1])
