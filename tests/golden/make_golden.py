"""
Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF
(/root/reference, unmodified) in the build container.

    PYTHONPATH=/root/reference python tests/golden/make_golden.py

The reference cannot travel to the GPU box, the fixtures can.  Inputs are rounded
to float32-representable values and handed to the reference as float64 arrays, so
the CUDA path and the reference see identical coordinates.

Files written (np.savez_compressed):
  golden_small.npz       1500 points, query == search, 3 scales (r/e = 5): reference output of
                         process_single_core plus, per scale, the reference VoxelFilter's unique
                         addresses/centres and the query_ball_tree neighbor lists (CSR).
  golden_urban.npz       2500-point surface-like scene, 5 scales with r/e = 3 (BASELINE config-2 shape),
                         a separate 800-point query cloud.
  golden_config1.npz     BASELINE config 1 (100k points, seed 10): the reference output for 400 of the
                         queries against the FULL 100k search cloud (the cloud is regenerated from its
                         seed by the tests).
  golden_degenerate.npz  queries with 0 / 1 / coincident neighbors.  the reference raises
                         FloatingPointError for those under numpy>=2 (features.py:43-55) although it
                         documents zeros (multiscale.py:4-5); here features.pca is wrapped to return
                         zeros in exactly those cases, everything else is the reference's own code.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
from nimrud.minimal import features as ref_features       # noqa: E402
from nimrud.minimal import multiscale as ref_multiscale   # noqa: E402
from nimrud.utils import geometry as ref_geometry         # noqa: E402
from scipy.spatial import cKDTree                          # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def f32(a):
    return np.asarray(a, dtype=np.float32).astype(np.float64)


def per_scale_internals(query, search, edge, radius):
    vf = ref_geometry.VoxelFilter(search, edge)
    addresses = np.unique(vf.coordinate_to_address(search))
    centres = vf.address_to_coordinate(addresses)
    tree = cKDTree(centres, leafsize=ref_multiscale.LEAFSIZE)
    lists = cKDTree(query, leafsize=ref_multiscale.LEAFSIZE).query_ball_tree(tree, radius)
    counts = np.array([len(l) for l in lists], dtype=np.int64)
    offsets = np.zeros(len(lists) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    flat = np.array([i for l in lists for i in sorted(l)], dtype=np.int32)
    return dict(min_corner=vf.minimum_corner, max_corner=vf.maximum_corner, widths=vf.widths,
                shifts=vf.shifts, addresses=addresses, centres=centres, offsets=offsets, indices=flat)


def pack(prefix, d):
    return {prefix + k: v for k, v in d.items()}


def make_small():
    rs = np.random.RandomState(1234)
    cloud = f32(rs.rand(1500, 3) * [4.0, 4.0, 1.5])
    edges = (0.1, 0.2, 0.4)
    radii = (0.5, 1.0, 2.0)
    feats = ref_multiscale.process_single_core(cloud, cloud, edges, radii)
    out = dict(query=cloud.astype(np.float32), search=cloud.astype(np.float32),
               edges=np.array(edges), radii=np.array(radii), features=feats)
    for s, (e, r) in enumerate(zip(edges, radii)):
        out.update(pack("s%d_" % s, per_scale_internals(cloud, cloud, e, r)))
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)
    print("small", feats.shape, feats[:, ::4].mean(0))


def urban_like(rs, n):
    """ground sheet + two boxes + a pole + a blob, about 12 x 12 x 6 m."""
    parts = []
    ng = int(n * 0.5)
    xy = rs.rand(ng, 2) * 12.0
    parts.append(np.c_[xy, 0.3 * np.sin(xy[:, 0] / 2.0) * np.cos(xy[:, 1] / 3.0) + rs.randn(ng) * 0.02])
    nb = int(n * 0.3)
    u = rs.rand(nb, 2)
    face = rs.randint(0, 5, nb)
    box = np.zeros((nb, 3))
    lo = np.array([3.0, 3.0, 0.0]); sz = np.array([4.0, 3.0, 5.0])
    box[:, 0] = np.where(face == 0, 0.0, np.where(face == 1, 1.0, u[:, 0]))
    box[:, 1] = np.where(face == 2, 0.0, np.where(face == 3, 1.0, np.where(face < 2, u[:, 0], u[:, 1])))
    box[:, 2] = np.where(face == 4, 1.0, np.where(face < 2, u[:, 1], u[:, 1]))
    parts.append(lo + box * sz)
    npole = int(n * 0.05)
    parts.append(np.c_[9.0 + rs.randn(npole) * 0.03, 9.0 + rs.randn(npole) * 0.03, rs.rand(npole) * 6.0])
    nv = n - ng - nb - npole
    parts.append(np.array([9.0, 3.0, 3.0]) + rs.randn(nv, 3) * 0.8)
    return f32(np.concatenate(parts))


def make_urban():
    rs = np.random.RandomState(20)
    search = urban_like(rs, 2500)
    query = search[rs.permutation(len(search))[:1200]].copy()
    edges = (0.1, 0.2, 0.4, 0.8, 1.6)
    radii = (0.3, 0.6, 1.2, 2.4, 4.8)
    # the reference as written raises when a neighborhood holds < 2 voxels (features.py:43-55):
    # keep the queries that see at least 2 voxels at every scale, so the UNMODIFIED reference runs.
    keep = np.ones(len(query), dtype=bool)
    for e, r in zip(edges, radii):
        keep &= np.diff(per_scale_internals(query, search, e, r)["offsets"]) >= 2
    query = query[keep][:800]
    out = dict(query=query.astype(np.float32), search=search.astype(np.float32),
               edges=np.array(edges), radii=np.array(radii))
    feats = ref_multiscale.process_single_core(query, search, edges, radii)
    out["features"] = feats
    for s, (e, r) in enumerate(zip(edges, radii)):
        out.update(pack("s%d_" % s, per_scale_internals(query, search, e, r)))
    np.savez_compressed(os.path.join(HERE, "golden_urban.npz"), **out)
    print("urban", feats.shape, feats[:, ::4].mean(0))


def make_config1():
    rs = np.random.RandomState(10)
    cloud = f32(rs.rand(100_000, 3) * [20, 20, 2])
    pick = np.random.RandomState(11).permutation(100_000)[:400]
    pick.sort()
    edges = (0.1, 0.2, 0.4)
    radii = (0.5, 1.0, 2.0)
    feats = ref_multiscale.process_single_core(cloud[pick], cloud, edges, radii)
    nvox = [len(np.unique(ref_geometry.VoxelFilter(cloud, e).coordinate_to_address(cloud))) for e in edges]
    np.savez_compressed(os.path.join(HERE, "golden_config1.npz"), pick=pick, edges=np.array(edges),
                        radii=np.array(radii), features=feats, n_voxels=np.array(nvox))
    print("config1", feats.shape, nvox, feats[:, ::4].mean(0))


def make_degenerate():
    rs = np.random.RandomState(5)
    search = f32(np.concatenate([
        rs.rand(400, 3) * [4.0, 4.0, 1.0],
        [[10.0, 10.0, 10.0]],                      # isolated voxel
        [[20.0, 0.0, 0.0], [20.01, 0.0, 0.0]],     # two raw points, one voxel
    ]))
    query = f32(np.concatenate([
        search[:50],
        [[10.0, 10.0, 10.0]],        # exactly one neighbor (its own voxel)
        [[10.2, 10.0, 10.0]],        # one neighbor, not at the voxel centre
        [[15.0, 15.0, 15.0]],        # no neighbors, inside the grid
        [[-30.0, -30.0, -30.0]],     # no neighbors, outside the grid
        [[20.0, 0.0, 0.0]],          # one voxel (two coincident-after-filter points)
    ]))
    edges = (0.2, 0.5)
    radii = (0.6, 1.0)

    stock_pca = ref_features.pca

    def guarded_pca(neighborhood_points):
        try:
            return stock_pca(neighborhood_points)
        except FloatingPointError:
            return np.zeros(2)

    ref_features.pca = guarded_pca
    try:
        feats = ref_multiscale.process_single_core(query, search, edges, radii)
    finally:
        ref_features.pca = stock_pca
    out = dict(query=query.astype(np.float32), search=search.astype(np.float32),
               edges=np.array(edges), radii=np.array(radii), features=feats)
    for s, (e, r) in enumerate(zip(edges, radii)):
        out.update(pack("s%d_" % s, per_scale_internals(query, search, e, r)))
    np.savez_compressed(os.path.join(HERE, "golden_degenerate.npz"), **out)
    print("degenerate", feats.shape)
    print(feats[50:])


if __name__ == "__main__":
    import warnings
    warnings.filterwarnings("ignore")
    make_small()
    make_urban()
    make_degenerate()
    make_config1()
