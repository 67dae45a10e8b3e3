"""
CPU oracle for the multiscale neighborhood eigenfeature path of grayhem/nimrud.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  Nothing under ``nimrud_b200/`` imports
``oracle``; the product path fails loudly when the CUDA library is missing.

It restates, function by function, what the reference computes (citations are
relative to /root/reference):

    voxel grid parameters      nimrud/utils/geometry.py:23-79   (VoxelFilter.__init__,
                                                                 _calculate_shift, _calculate_masks)
    coordinate -> address      nimrud/utils/geometry.py:103-116
    address -> coordinate      nimrud/utils/geometry.py:120-138
    unique voxels              nimrud/utils/geometry.py:142-154
    radius neighbor sets       nimrud/minimal/multiscale.py:87,100,103  (scipy cKDTree,
                               third-party, unpinned by the reference; scipy 1.18.1 here)
    population/centroid/pca    nimrud/minimal/features.py:14-57
    one scale / all scales     nimrud/minimal/multiscale.py:27-123

Parity status:
  * voxel filter: PINNED by the reference's known-answer tests
    (nimrud/utils/tests/geometry_tests.py:84-279), re-checked in tests/test_oracle.py.
  * radius sets / population / centroid / eigen ratios: the reference holds no
    tests for these; pinned here against outputs of the reference itself, run in
    the build container and committed as tests/golden/*.npz by
    tests/golden/make_golden.py.
  * kNN and extended descriptors: no reference code exists; PARITY UNPINNED.
    The oracle below is the definition (brute force, (d^2, index) order;
    numpy eigh on the oracle's own covariance).

The arithmetic that decides neighbor membership is reproduced literally:
float64, inclusive, ``dx*dx + dy*dy + dz*dz <= r*r`` summed left to right, on
voxel centres reconstructed as ``(k*e + min_corner) + e*0.5``.
"""

import numpy as np

MAX_ADDRESS_LENGTH = 64   # geometry.py:13


# --------------------------------------------------------------------------------------
# voxel filter (a1-a4)
# --------------------------------------------------------------------------------------

class GridParams(object):
    """the numbers VoxelFilter.__init__ derives from a cloud (geometry.py:37-48)."""

    def __init__(self, minimum_corner, maximum_corner, edge_length):
        self.minimum_corner = np.asarray(minimum_corner, dtype=np.float64)
        self.maximum_corner = np.asarray(maximum_corner, dtype=np.float64)
        self.edge_length = edge_length
        extent = self.maximum_corner - self.minimum_corner
        # geometry.py:55-56
        bits = np.ceil(np.log2(extent / edge_length))
        if bits.sum() > MAX_ADDRESS_LENGTH:
            # geometry.py:59-60
            raise ValueError("edge length is too small to address this space")
        self.widths = bits.astype(np.int64)
        self.shifts = np.cumsum(bits)[:-1].astype(np.int64)          # geometry.py:62
        # geometry.py:74-77 : width ones, moved up by the shift of that axis
        offsets = np.concatenate(([0], self.shifts))
        self.masks = [((1 << int(w)) - 1) << int(s) for w, s in zip(self.widths, offsets)]


def grid_params(points, edge_length):
    """a1: geometry.py:23-48 including the shape validation at :30-35."""
    points = np.asarray(points)
    if points.ndim != 2:
        raise ValueError("wrong point cloud array shape")
    if points.shape[1] not in (2, 3):
        raise ValueError("only 2D and 3D spaces supported")
    if points.shape[0] < 2:
        raise ValueError("need at least 2 points to define a voxel grid")
    points = points.astype(np.float64, copy=False)
    half = edge_length / 2
    return GridParams(points.min(0) - half, points.max(0) + half, edge_length)


def coordinate_to_address(params, points):
    """a2: geometry.py:103-116 (bounds check geometry.py:83-99)."""
    pts = np.atleast_2d(np.asarray(points, dtype=np.float64))
    if pts.ndim != 2:
        raise ValueError("wrong array shape")
    if pts.shape[1] != params.shifts.size + 1:
        raise ValueError("wrong number of spatial dimensions")
    if (pts.min(0) < params.minimum_corner).any() or (pts.max(0) > params.maximum_corner).any():
        raise ValueError("some points fall outside filter bounding region")
    cells = np.floor((pts - params.minimum_corner) / params.edge_length).astype(np.int64)
    address = cells[:, 0].copy()
    for axis, shift in enumerate(params.shifts):
        address += cells[:, axis + 1] << shift
    return address


def address_to_cells(params, addresses):
    addresses = np.atleast_1d(np.asarray(addresses, dtype=np.int64))
    offsets = np.concatenate(([0], params.shifts))
    return np.stack([(addresses & m) >> int(s) for m, s in zip(params.masks, offsets)], axis=1)


def address_to_coordinate(params, addresses):
    """a4: geometry.py:120-138.  operation order matters: (k*e + min) + e*0.5."""
    cells = address_to_cells(params, addresses)
    return cells * params.edge_length + params.minimum_corner + params.edge_length * 0.5


def unique_voxels(params, points):
    """a3: geometry.py:142-154; returns (sorted unique addresses, centres)."""
    keys = np.unique(coordinate_to_address(params, points))
    return keys, address_to_coordinate(params, keys)


# --------------------------------------------------------------------------------------
# neighbor sets (a5)
# --------------------------------------------------------------------------------------

def radius_sets(query, centres, radius, chunk=1000, leafsize=300):
    """
    a5: multiscale.py:87,100,103.  same scipy call the reference makes (chunk tree
    against search tree); every list is returned sorted ascending.  returned as CSR
    (offsets int64 (Nq+1,), indices int64).
    """
    from scipy.spatial import cKDTree
    search_tree = cKDTree(centres, leafsize=leafsize)
    counts = []
    flat = []
    for start in range(0, len(query), chunk):
        block = query[start:start + chunk]
        lists = cKDTree(block, leafsize=leafsize).query_ball_tree(search_tree, radius)
        for one in lists:
            one = sorted(one)
            counts.append(len(one))
            flat.extend(one)
    offsets = np.zeros(len(query) + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    return offsets, np.asarray(flat, dtype=np.int64)


def radius_sets_bruteforce(query, centres, radius):
    """
    literal statement of the membership predicate, independent of scipy's tree:
    s = dx*dx; s += dy*dy; s += dz*dz; member iff s <= r*r   (float64, no fma).
    O(Nq*Nv): small cases only.  used to pin `radius_sets` itself.
    """
    r2 = radius * radius
    offsets = [0]
    flat = []
    for q in query:
        d = q[None, :] - centres
        s = d[:, 0] * d[:, 0]
        s = s + d[:, 1] * d[:, 1]
        s = s + d[:, 2] * d[:, 2]
        idx = np.nonzero(s <= r2)[0]
        flat.append(idx)
        offsets.append(offsets[-1] + idx.size)
    flat = np.concatenate(flat) if flat else np.zeros(0, dtype=np.int64)
    return np.asarray(offsets, dtype=np.int64), flat.astype(np.int64)


def knn_sets(query, points, k):
    """
    kNN oracle (no reference code; PARITY UNPINNED): brute force, float64 squared
    distance summed x,y,z left to right, total order (d^2, index).
    returns (indices (Nq,k) int64, d2 (Nq,k) float64); rows padded with -1/inf if
    fewer than k points exist.
    """
    nq = len(query)
    idx_out = np.full((nq, k), -1, dtype=np.int64)
    d2_out = np.full((nq, k), np.inf, dtype=np.float64)
    index = np.arange(len(points))
    for i, q in enumerate(query):
        d = q[None, :] - points
        s = d[:, 0] * d[:, 0]
        s = s + d[:, 1] * d[:, 1]
        s = s + d[:, 2] * d[:, 2]
        order = np.lexsort((index, s))[:k]
        idx_out[i, :order.size] = order
        d2_out[i, :order.size] = s[order]
    return idx_out, d2_out


# --------------------------------------------------------------------------------------
# per-neighborhood features (a6-a9)
# --------------------------------------------------------------------------------------

def neighborhood_row(query_point, neighbors):
    """
    a7,a8,a9 for one neighborhood -> [population, centroid, l_max/sum, l_mid/sum].
    features.py:21-57 with the module's DOCUMENTED behaviour for undefined features
    (multiscale.py:4-5: zeros) where the code as written raises under numpy>=2
    (n<2 or zero total variance).
    """
    n = neighbors.shape[0]
    if n == 0:
        return np.zeros(4)
    centroid = np.linalg.norm(query_point - neighbors.mean(0))          # features.py:26
    ratios = np.zeros(2)
    if n >= 2:
        eigvals = np.linalg.eigvalsh(np.cov(neighbors, rowvar=False))   # features.py:43,46
        total = eigvals.sum()
        if total != 0:
            eigvals = eigvals / total                                   # features.py:55
            ratios = eigvals[:0:-1]                                     # features.py:57
    return np.array([n, centroid, ratios[0], ratios[1]])


def extended_row(neighbors):
    """
    extension descriptors (no reference code; PARITY UNPINNED).  l1>=l2>=l3 of the
    same covariance, e_i = l_i/sum:
      linearity (e1-e2)/e1, planarity (e2-e3)/e1, sphericity e3/e1,
      omnivariance (e1 e2 e3)^(1/3), anisotropy (e1-e3)/e1,
      eigenentropy -sum e_i ln e_i, change of curvature e3,
      normal (unit eigenvector of l3, sign fixed to nz>=0), verticality 1-|nz|.
    returns 22 numbers: the 8 scalars, nx, ny, nz, l-sum, the upper triangle of the covariance
    (xx xy xz yy yz zz, numpy.cov, ddof = 1), and x, y of the unit eigenvectors of l1 and l2 (sign: x > 0, else
    y > 0, else z > 0; the legacy OG_MSO keeps two components of two eigenvectors, prototypes/mso.py:1498-1539).
    zeros if undefined.
    """
    out = np.zeros(22)
    if neighbors.shape[0] < 3:
        return out
    cov = np.cov(neighbors, rowvar=False)
    w, v = np.linalg.eigh(cov)
    total = w.sum()
    if not total > 0:
        return out
    w = np.clip(w, 0.0, None)
    e3, e2, e1 = w / total
    normal = v[:, 0]
    if normal[2] < 0 or (normal[2] == 0 and (normal[1] < 0 or (normal[1] == 0 and normal[0] < 0))):
        normal = -normal
    ent = 0.0
    for e in (e1, e2, e3):
        if e > 0:
            ent -= e * np.log(e)
    def canonical(w):
        if w[0] < 0 or (w[0] == 0 and (w[1] < 0 or (w[1] == 0 and w[2] < 0))):
            return -w
        return w
    lead, second = canonical(v[:, 2]), canonical(v[:, 1])
    out[:] = [(e1 - e2) / e1, (e2 - e3) / e1, e3 / e1, np.cbrt(e1 * e2 * e3), (e1 - e3) / e1,
              ent, e3, 1.0 - abs(normal[2]), normal[0], normal[1], normal[2], total,
              cov[0, 0], cov[0, 1], cov[0, 2], cov[1, 1], cov[1, 2], cov[2, 2],
              lead[0], lead[1], second[0], second[1]]
    return out


def rows_from_sets(query, centres, offsets, indices, row_fn=None):
    nq = len(query)
    if row_fn is None:
        out = np.zeros((nq, 4))
        for i in range(nq):
            out[i] = neighborhood_row(query[i], centres[indices[offsets[i]:offsets[i + 1]]])
    else:
        first = row_fn(centres[:0])
        out = np.zeros((nq, first.size))
        for i in range(nq):
            out[i] = row_fn(centres[indices[offsets[i]:offsets[i + 1]]])
    return out


# --------------------------------------------------------------------------------------
# drivers (a10, a11)
# --------------------------------------------------------------------------------------

def one_scale(query, search, edge_length, radius):
    """a10: multiscale.py:70-123 -> (Nq,4) float64."""
    query = np.asarray(query, dtype=np.float64)
    search = np.asarray(search, dtype=np.float64)
    params = grid_params(search, edge_length)
    _, centres = unique_voxels(params, search)
    offsets, indices = radius_sets(query, centres, radius)
    return rows_from_sets(query, centres, offsets, indices)


def process(query, search, edge_lengths, radii):
    """a11: multiscale.py:27-67 -> (Nq, 4*S) float64, scale-major columns."""
    assert len(edge_lengths) == len(radii), \
        "edge_lengths and radii should be equal-length sequences."
    return np.concatenate(
        [one_scale(query, search, e, r) for e, r in zip(edge_lengths, radii)], axis=1)


def knn_features(query, points, ks):
    """kNN 'scales': the 4 reference columns over the k nearest, for each k. (Nq, 4*len(ks))"""
    query = np.asarray(query, dtype=np.float64)
    points = np.asarray(points, dtype=np.float64)
    kmax = max(ks)
    idx, _ = knn_sets(query, points, kmax)
    blocks = []
    for k in ks:
        block = np.zeros((len(query), 4))
        for i in range(len(query)):
            sel = idx[i, :k]
            sel = sel[sel >= 0]
            block[i] = neighborhood_row(query[i], points[sel])
        blocks.append(block)
    return np.concatenate(blocks, axis=1)


def _shard(args):
    query, search, edges, radii = args
    return process(query, search, edges, radii)


def process_parallel(query, search, edge_lengths, radii, workers):
    """
    the parallelisation the reference's author proposes (multiscale.py:92-94): query shards in
    a multiprocessing pool.  the result is bitwise independent of the sharding.
    """
    import multiprocessing as mp
    if workers <= 1:
        return process(query, search, edge_lengths, radii)
    pieces = np.array_split(np.asarray(query, dtype=np.float64), workers)
    with mp.get_context("fork").Pool(workers) as pool:
        blocks = pool.map(_shard, [(p, search, tuple(edge_lengths), tuple(radii)) for p in pieces])
    return np.concatenate(blocks, axis=0)
