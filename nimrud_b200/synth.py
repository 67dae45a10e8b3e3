"""
synthetic point clouds for the BASELINE.json configs.  torch generators so that the
10M / 100M clouds can be produced directly on the GPU; all outputs float32, shape (n, 3).

These are workload generators for tests and benches, not part of the reference's API.
"""
import math

import numpy as np
import torch


def uniform_box(n=100_000, extent=(20.0, 20.0, 2.0), seed=10):
    """BASELINE config 1: np.random.RandomState(seed).rand(n,3) * extent, rounded to float32."""
    rs = np.random.RandomState(seed)
    return (rs.rand(n, 3) * np.asarray(extent)).astype(np.float32)


def _gen(seed, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def urban_scene(n, seed=20, density=40.0, device="cpu", origin=(0.0, 0.0), return_labels=False):
    """
    BASELINE config 2 / 5: urban scene at ~`density` points per square metre of ground footprint.
    45 % ground (gentle relief + 2 cm noise), 30 % axis-aligned box buildings (walls + roofs),
    5 % poles and wires, 20 % vegetation blobs.  extent = sqrt(n / density) (500 m for n = 10M).
    labels: 0 ground, 1 building, 2 pole/wire, 3 vegetation.
    """
    device = torch.device(device)
    g = _gen(seed, device)
    extent = math.sqrt(n / density)

    def rand(*shape):
        return torch.rand(*shape, generator=g, device=device, dtype=torch.float32)

    def randn(*shape):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.float32)

    n_ground = int(n * 0.45)
    n_build = int(n * 0.30)
    n_pole = int(n * 0.05)
    n_veg = n - n_ground - n_build - n_pole
    parts, labels = [], []

    # ground
    xy = rand(n_ground, 2) * extent
    z = 0.5 * torch.sin(xy[:, 0] / 40.0) * torch.cos(xy[:, 1] / 55.0) + randn(n_ground) * 0.02
    parts.append(torch.cat([xy, z[:, None]], 1)); labels.append(torch.zeros(n_ground, dtype=torch.int8, device=device))

    # buildings: ~200 boxes per 500 m x 500 m
    n_boxes = max(2, int(round(200 * (extent / 500.0) ** 2)))
    foot = 8.0 + rand(n_boxes, 2) * 22.0
    height = 6.0 + rand(n_boxes) * 34.0
    corner = rand(n_boxes, 2) * max(extent - 30.0, 1.0)
    which = torch.randint(0, n_boxes, (n_build,), generator=g, device=device)
    face = torch.randint(0, 5, (n_build,), generator=g, device=device)
    u = rand(n_build, 2)
    fx, fy, h = foot[which, 0], foot[which, 1], height[which]
    bx = torch.where(face == 0, torch.zeros_like(fx), torch.where(face == 1, fx, u[:, 0] * fx))
    by = torch.where(face == 2, torch.zeros_like(fy), torch.where(face == 3, fy,
                     torch.where(face < 2, u[:, 0] * fy, u[:, 1] * fy)))
    bz = torch.where(face == 4, h, u[:, 1] * h)
    parts.append(torch.stack([corner[which, 0] + bx, corner[which, 1] + by, bz], 1))
    labels.append(torch.ones(n_build, dtype=torch.int8, device=device))

    # poles (vertical, h = 8 m, r = 5 cm) and wires (horizontal at z = 7)
    n_poles = max(2, int(round(400 * (extent / 500.0) ** 2)))
    pole_xy = rand(n_poles, 2) * extent
    n_p = n_pole // 2
    wp = torch.randint(0, n_poles, (n_p,), generator=g, device=device)
    parts.append(torch.cat([pole_xy[wp] + randn(n_p, 2) * 0.05, rand(n_p, 1) * 8.0], 1))
    n_w = n_pole - n_p
    ww = torch.randint(0, n_poles - 1, (n_w,), generator=g, device=device)
    t = rand(n_w, 1)
    wire_xy = pole_xy[ww] * (1 - t) + pole_xy[ww + 1] * t
    parts.append(torch.cat([wire_xy, 7.0 + randn(n_w, 1) * 0.01], 1))
    labels.append(torch.full((n_pole,), 2, dtype=torch.int8, device=device))

    # vegetation blobs
    n_blobs = max(2, int(round(1500 * (extent / 500.0) ** 2)))
    centre = torch.cat([rand(n_blobs, 2) * extent, 3.0 + rand(n_blobs, 1) * 7.0], 1)
    sigma = 1.5 + rand(n_blobs) * 1.5
    wb = torch.randint(0, n_blobs, (n_veg,), generator=g, device=device)
    parts.append(centre[wb] + randn(n_veg, 3) * sigma[wb, None])
    labels.append(torch.full((n_veg,), 3, dtype=torch.int8, device=device))

    cloud = torch.cat(parts, 0)
    lab = torch.cat(labels, 0)
    perm = torch.randperm(n, generator=g, device=device)
    cloud = cloud[perm].contiguous()
    cloud[:, 0] += origin[0]
    cloud[:, 1] += origin[1]
    if return_labels:
        return cloud, lab[perm].contiguous()
    return cloud


def aerial_tile(n, seed=22, density=8.0, device="cpu", origin=(0.0, 0.0)):
    """
    BASELINE config 4: aerial-LiDAR-like tile at ~`density` points per square metre.  60 % terrain
    (4 sinusoid octaves, amplitude 30 m), 15 % building roofs, 25 % canopy blobs.
    the terrain is a function of the GLOBAL coordinate (origin + local), so tiles generated with
    different origins join seamlessly.  extent = sqrt(n / density) (3.5 km for n = 100M).
    """
    device = torch.device(device)
    g = _gen(seed, device)
    extent = math.sqrt(n / density)

    def rand(*shape):
        return torch.rand(*shape, generator=g, device=device, dtype=torch.float32)

    def randn(*shape):
        return torch.randn(*shape, generator=g, device=device, dtype=torch.float32)

    def terrain(x, y):
        z = torch.zeros_like(x)
        amp, wav = 30.0, 900.0
        for _ in range(4):
            z = z + amp * torch.sin(x / wav * 6.2831853) * torch.cos(y / (wav * 1.3) * 6.2831853)
            amp *= 0.45
            wav *= 0.4
        return z

    n_ter = int(n * 0.60)
    n_roof = int(n * 0.15)
    n_can = n - n_ter - n_roof
    parts = []
    xy = rand(n_ter, 2) * extent
    gx, gy = xy[:, 0] + origin[0], xy[:, 1] + origin[1]
    parts.append(torch.stack([gx, gy, terrain(gx, gy) + randn(n_ter) * 0.05], 1))

    n_roofs = max(2, int(round(n_roof / 2000)))
    rc = rand(n_roofs, 2) * extent
    rs_ = 6.0 + rand(n_roofs, 2) * 20.0
    rh = 4.0 + rand(n_roofs) * 12.0
    wr = torch.randint(0, n_roofs, (n_roof,), generator=g, device=device)
    u = rand(n_roof, 2)
    rx = rc[wr, 0] + u[:, 0] * rs_[wr, 0] + origin[0]
    ry = rc[wr, 1] + u[:, 1] * rs_[wr, 1] + origin[1]
    base = terrain(rc[wr, 0] + origin[0], rc[wr, 1] + origin[1])
    parts.append(torch.stack([rx, ry, base + rh[wr] + randn(n_roof) * 0.03], 1))

    n_blobs = max(2, int(round(n_can / 800)))
    bc = rand(n_blobs, 2) * extent
    sg = 1.5 + rand(n_blobs) * 2.5
    wb = torch.randint(0, n_blobs, (n_can,), generator=g, device=device)
    cx = bc[wb, 0] + randn(n_can) * sg[wb] + origin[0]
    cy = bc[wb, 1] + randn(n_can) * sg[wb] + origin[1]
    cz = terrain(bc[wb, 0] + origin[0], bc[wb, 1] + origin[1]) + 6.0 + randn(n_can) * sg[wb]
    parts.append(torch.stack([cx, cy, cz], 1))

    cloud = torch.cat(parts, 0)
    perm = torch.randperm(n, generator=g, device=device)
    return cloud[perm].contiguous()


def with_ties(cloud, edge_length, seed=21, fraction=0.01):
    """
    BASELINE config 3 query cloud: append `fraction` exact duplicates and `fraction` points placed
    exactly on lattice positions (multiples of edge_length), which forces equal-distance ties.
    """
    g = _gen(seed, cloud.device)
    n = cloud.shape[0]
    m = max(1, int(n * fraction))
    dup = cloud[torch.randint(0, n, (m,), generator=g, device=cloud.device)]
    pick = cloud[torch.randint(0, n, (m,), generator=g, device=cloud.device)]
    snapped = torch.round(pick / edge_length) * edge_length
    return torch.cat([cloud, dup, snapped.to(cloud.dtype)], 0).contiguous()
