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
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), feats.cpu().numpy())
    dist.barrier()
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
