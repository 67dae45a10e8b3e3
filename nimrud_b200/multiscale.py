"""
multiscale operator processing pipeline -- B200 drop-in for `nimrud.minimal.multiscale`
(reference: nimrud/minimal/multiscale.py).

features are generated for points in the query cloud, using geometry from the search cloud.
all undefined features are represented by zeros (reference: multiscale.py:4-5).

`process_single_core` / `one_scale_single_core` keep the reference's names, argument order and
output layout: (Nq, 4*S) float64, column 4*s+j = feature j of scale s, features
[population, centroid distance, l_max/sum, l_mid/sum].

numpy arrays in -> numpy array out (host buffers through nbr_multiscale_features_host);
CUDA torch tensors in -> CUDA torch tensor out (nbr_multiscale_features on the current stream).
There is no CPU implementation in this package.
"""
import ctypes
import time

import numpy as np
import torch

from . import _lib, _results
from ._util import device_cloud, host_cloud, is_torch, ptr, require_cuda, stream_ptr, validate_cloud
from .geometry import cloud_bbox, grid_from_bbox

# the reference's tunables (multiscale.py:18-24).  results never depended on them; the CUDA path
# has no leaf size and no query chunking, they are kept so that code that sets them still imports.
LEAFSIZE = 300
QUERY_CHUNK_SIZE = 1000
VERBOSITY_INTERVAL = 100

_TORCH_OUT = {np.float64: torch.float64, np.float32: torch.float32}


def _out_code(out_dtype):
    dt = np.dtype(out_dtype)
    if dt == np.float64:
        return np.float64, _lib.F64
    if dt == np.float32:
        return np.float32, _lib.F32
    raise ValueError("out_dtype must be float32 or float64")


def _ncol(descriptors):
    if descriptors in ("reference", 0, None):
        return 4, _lib.DESC_REFERENCE
    if descriptors in ("extended", 1):
        return 26, _lib.DESC_EXTENDED
    raise ValueError("descriptors must be 'reference' or 'extended'")


def _check_inputs(query_cloud, search_cloud):
    # reference: VoxelFilter.__init__ on the search cloud (utils/geometry.py:30-35); scipy rejects
    # a query cloud whose width differs from the tree's.
    if search_cloud.ndim != 2:
        raise ValueError("wrong point cloud array shape")
    if search_cloud.shape[1] not in (2, 3):
        raise ValueError("only 2D and 3D spaces supported")
    if search_cloud.shape[0] < 2:
        raise ValueError("need at least 2 points to define a voxel grid")
    if search_cloud.shape[1] != 3:
        raise ValueError("the eigenfeature path needs 3D clouds")
    validate_cloud(query_cloud, "query_cloud")


def process_single_core(query_cloud, search_cloud, edge_lengths, radii, verbose=False,
                        out_dtype=np.float64, descriptors="reference", return_voxel_counts=False,
                        global_bbox=None, out=None):
    """
    compute features at multiple scales. returns an array of feature vectors aligned with the query
    cloud.  (reference: multiscale.py:27-67)

    extras beyond the reference signature (keyword only in spirit, defaults = reference behaviour):
      out_dtype            float64 (drop-in) or float32
      descriptors          "reference" (4 columns per scale) or "extended" (26 columns per scale)
      return_voxel_counts  also return the number of unique search voxels per scale
      global_bbox          (lo, hi) of the WHOLE search cloud when `search_cloud` is only a tile of it
                           (+ halo): the voxel lattices are anchored on that box, so every tile sees the
                           same voxels (multi-GPU path, CUDA tensors only)
      out                  preallocated CUDA output tensor (CUDA tensors only)
    """
    assert(len(edge_lengths) == len(radii)), \
        "edge_lengths and radii should be equal-length sequences."
    _check_inputs(query_cloud, search_cloud)
    require_cuda()
    np_out, out_code = _out_code(out_dtype)
    ncol, mask = _ncol(descriptors)
    n_scales = len(radii)
    nq = int(query_cloud.shape[0])
    ns = int(search_cloud.shape[0])
    edges_arr, edges_p = _lib.f64_array(list(edge_lengths))
    radii_arr, radii_p = _lib.f64_array(list(radii))
    counts = np.zeros(max(n_scales, 1), dtype=np.int64)
    counts_p = counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)) if (return_voxel_counts or verbose) else None
    start = time.perf_counter()

    if is_torch(query_cloud) and query_cloud.is_cuda:
        q, qc = device_cloud(query_cloud)
        if search_cloud is query_cloud:
            s, sc = q, qc
        else:
            s, sc = device_cloud(search_cloud, q.device)
        if out is None:
            out = torch.zeros((nq, ncol * n_scales), dtype=_TORCH_OUT[np_out], device=q.device)
        elif (tuple(out.shape) != (nq, ncol * n_scales) or out.dtype != _TORCH_OUT[np_out] or not out.is_cuda
              or not out.is_contiguous()):
            raise ValueError("out has the wrong shape, dtype or device")
        box_p = None
        if global_bbox is not None:
            box_arr, box_p = _lib.f64_array(np.concatenate([np.asarray(global_bbox[0], dtype=np.float64),
                                                             np.asarray(global_bbox[1], dtype=np.float64)]))
        try:
            with torch.cuda.device(q.device):
                _lib.check(_lib.lib().nbr_multiscale_features(
                    ptr(q), qc, nq, ptr(s), sc, ns, edges_p, radii_p, n_scales, ptr(out), out_code, mask, box_p,
                    counts_p, stream_ptr(q.device)))
        except NotImplementedError as err:
            if "directory" not in str(err) or global_bbox is not None or counts_p is not None:
                raise
            _process_in_tiles(q, s, edge_lengths, radii, out, out_dtype, descriptors)
    else:
        if global_bbox is not None or out is not None:
            raise ValueError("global_bbox / out are only supported for CUDA tensor inputs")
        q, qc = host_cloud(query_cloud.cpu().numpy() if is_torch(query_cloud) else query_cloud)
        if search_cloud is query_cloud:
            s, sc = q, qc
        else:
            s, sc = host_cloud(search_cloud.cpu().numpy() if is_torch(search_cloud) else search_cloud)
        # every element is written by the call; an untouched allocation lets the host threads that fill it fault
        # its pages in parallel
        out = _results.empty((nq, ncol * n_scales), np_out)
        if n_scales == 0 or nq == 0:
            out[...] = 0
        rc = _lib.lib().nbr_multiscale_features_host(
            ctypes.c_void_p(q.ctypes.data), qc, nq, ctypes.c_void_p(s.ctypes.data), sc, ns, edges_p, radii_p,
            n_scales, ctypes.c_void_p(out.ctypes.data), out_code, mask, counts_p)
        if rc == _lib.ERR_UNSUPPORTED and "directory" in _lib.last_error() and counts_p is None:
            # sparse cloud over a large extent: tile by tile on the device (see _process_in_tiles)
            dq = torch.from_numpy(q).cuda()
            ds = dq if search_cloud is query_cloud else torch.from_numpy(s).cuda()
            dout = torch.zeros((nq, ncol * n_scales), dtype=_TORCH_OUT[np_out], device=dq.device)
            _process_in_tiles(dq, ds, edge_lengths, radii, dout, out_dtype, descriptors)
            out[...] = dout.cpu().numpy()
        else:
            _lib.check(rc)
        if is_torch(query_cloud):
            out = torch.from_numpy(out)

    if verbose:
        if is_torch(out) and out.is_cuda:
            torch.cuda.synchronize(out.device)
        elapsed = time.perf_counter() - start
        for e, r, c in zip(edges_arr, radii_arr, counts):
            print("queried {} points against a search space of {} voxels".format(nq, int(c)))
            print("using a voxel edge length of {} and radius of {}".format(e, r))
        print("===============================")
        print("calculating all scales took {}s".format(np.around(elapsed, 3)))
        print("final rate of {} points per second".format(np.around(nq / max(elapsed, 1e-12), 3)))
    if return_voxel_counts:
        return out, counts[:n_scales].copy()
    return out


def _process_in_tiles(q, s, edge_lengths, radii, out, out_dtype, descriptors, max_dir_entries=4.0e8):
    """
    a cloud whose bounding box is too large for a dense brick directory (the reference only needs the packed voxel
    address to fit 64 bits, utils/geometry.py:55-60): the x-y plane is cut into square tiles small enough for their
    directories, every tile's queries run against the search points of the tile grown by h = max(r + e/2), and
    every lattice stays anchored on the WHOLE search cloud's box -- the same voxels, hence the same rows, as one
    unpartitioned call (the multi-GPU tile path proves the same decomposition bit for bit, tests/test_tiles_gpu.py).
    q, s: CUDA tensors; out: preallocated CUDA tensor, rows in the queries' order.
    """
    from .distributed import halo_width
    ncol, mask = _ncol(descriptors)
    np_out, out_code = _out_code(out_dtype)
    lo = s.min(0).values.double()
    hi = s.max(0).values.double()
    g_lo, g_hi = lo.cpu().numpy(), hi.cpu().numpy()
    grid_from_bbox(g_lo, g_hi, float(min(edge_lengths)), 3)          # ValueError if the address needs more than 64 bits
    h = halo_width(edge_lengths, radii)
    e_min = float(min(edge_lengths))
    lz = float(g_hi[2] - g_lo[2]) + 2 * h + 8 * e_min
    # directory entries of a tile of side T: (T + 2h)^2 * lz / (1024 e^3) for the finest lattice (bricks of 32 x 8 x 4)
    side = (max_dir_entries * 1024.0 * e_min ** 3 / lz) ** 0.5 - 2 * h
    if not side > 4 * h:
        raise NotImplementedError("cloud too tall / edge too small for tiled processing (tile side %.3g, halo %.3g)" % (side, h))
    nx = max(1, int(np.ceil((g_hi[0] - g_lo[0]) / side)))
    ny = max(1, int(np.ceil((g_hi[1] - g_lo[1]) / side)))
    if nx * ny > 1 << 20:
        raise NotImplementedError("cloud extent needs more than 2^20 tiles")

    def tile_ids(c):
        ix = ((c[:, 0].double() - lo[0]) / side).floor().clamp_(0, nx - 1).long()
        iy = ((c[:, 1].double() - lo[1]) / side).floor().clamp_(0, ny - 1).long()
        return iy * nx + ix

    q_tile = tile_ids(q)
    q_order = torch.argsort(q_tile)
    q_sorted_tile = q_tile[q_order]
    tiles, q_counts = torch.unique_consecutive(q_sorted_tile, return_counts=True)
    q_starts = torch.cumsum(q_counts, 0) - q_counts
    same = s is q
    s_tile = q_tile if same else tile_ids(s)
    s_order = q_order if same else torch.argsort(s_tile)
    s_sorted_tile = s_tile[s_order]
    s_starts_all = torch.searchsorted(s_sorted_tile, torch.arange(nx * ny + 1, device=s.device))
    s_starts_all = s_starts_all.cpu().numpy()
    out.zero_()
    for t, q0, qn in zip(tiles.cpu().tolist(), q_starts.cpu().tolist(), q_counts.cpu().tolist()):
        ty, tx = divmod(t, nx)
        rows = q_order[q0:q0 + qn]
        t_lo = torch.tensor([float(g_lo[0]) + tx * side - h, float(g_lo[1]) + ty * side - h], dtype=torch.float64, device=s.device)
        t_hi = t_lo + side + 2 * h
        parts = []
        for yy in range(max(ty - 1, 0), min(ty + 1, ny - 1) + 1):
            for xx in range(max(tx - 1, 0), min(tx + 1, nx - 1) + 1):
                a, b = int(s_starts_all[yy * nx + xx]), int(s_starts_all[yy * nx + xx + 1])
                if b > a:
                    cand = s[s_order[a:b]]
                    if (yy, xx) != (ty, tx):
                        keep = ((cand[:, :2].double() >= t_lo) & (cand[:, :2].double() <= t_hi)).all(1)
                        cand = cand[keep]
                    parts.append(cand)
        if not parts:
            continue
        search = torch.cat(parts, 0).contiguous() if len(parts) > 1 else parts[0].contiguous()
        if search.shape[0] < 2:
            search = torch.cat([search, search], 0)                    # one point: the same voxel set, and the C entry wants two
        tile_out = process_single_core(q[rows].contiguous(), search, edge_lengths, radii, out_dtype=out_dtype,
                                       descriptors=descriptors, global_bbox=(g_lo, g_hi))
        out[rows] = tile_out
    return out


def one_scale_single_core(query_cloud, search_cloud, edge_length, radius, verbose=False, **kwargs):
    """
    generate a 4d feature vector representing one analysis scale.  (reference: multiscale.py:70-123)
    """
    return process_single_core(query_cloud, search_cloud, [edge_length], [radius], verbose=verbose, **kwargs)


# --------------------------------------------------------------------------------------------------
# the pieces, for callers that keep an index around (and for the parity tests)
# --------------------------------------------------------------------------------------------------

class LatticeIndex(object):
    """
    the voxel-filtered search cloud of one edge length on the GPU (what the reference builds at
    multiscale.py:75-87: VoxelFilter.unique_voxels + cKDTree), as occupancy bit bricks.

    indexed=True additionally keeps the sorted unique addresses so neighbor INDICES (positions in
    np.unique order, as cKDTree reports them) can be returned.
    """

    def __init__(self, search_cloud, edge_length, indexed=False, bbox=None):
        if search_cloud.ndim != 2:
            raise ValueError("wrong point cloud array shape")
        if search_cloud.shape[1] != 3:
            raise ValueError("only 3D search clouds can be indexed")
        if search_cloud.shape[0] < 2:
            raise ValueError("need at least 2 points to define a voxel grid")
        require_cuda()
        self._search, self._code = device_cloud(search_cloud)
        self.device = self._search.device
        self.edge_length = edge_length
        lo, hi = bbox if bbox is not None else cloud_bbox(self._search, self._code)
        self.grid = grid_from_bbox(lo, hi, edge_length, 3)
        self.indexed = bool(indexed)
        handle = ctypes.c_void_p()
        # the lattice is built (and later freed) on the stream that is current now; calls made under another
        # stream are ordered after the build, and the release after them (_sync_streams)
        self._stream = torch.cuda.current_stream(self.device)
        self._used_on = set()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_lattice_create(
                ctypes.byref(handle), ptr(self._search), self._code, self._search.shape[0], ctypes.byref(self.grid),
                _lib.LATTICE_INDEXED if indexed else 0, stream_ptr(self.device)))
        self._handle = handle
        self._counts = None

    def _sync_streams(self):
        """order the current stream after the lattice's build stream (no-op when they are the same stream)."""
        cur = torch.cuda.current_stream(self.device)
        if cur != self._stream:
            cur.wait_stream(self._stream)
            self._used_on.add(cur)

    def close(self):
        if getattr(self, "_handle", None):
            with torch.cuda.device(self.device):       # frees are ordered on the build stream: after every user
                for st in getattr(self, "_used_on", ()):
                    self._stream.wait_stream(st)
                _lib.lib().nbr_lattice_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _info(self):
        if self._counts is None:
            nv, nb = ctypes.c_int64(), ctypes.c_int64()
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().nbr_lattice_info(self._handle, ctypes.byref(nv), ctypes.byref(nb)))
            self._counts = (nv.value, nb.value)
        return self._counts

    @property
    def n_voxels(self):
        return self._info()[0]

    @property
    def n_bricks(self):
        return self._info()[1]

    def addresses_and_centres(self):
        """(sorted unique addresses int64 (Nv,), centres float64 (Nv,3)) as CUDA tensors."""
        self._sync_streams()
        nv = self.n_voxels
        addr = torch.empty(nv, dtype=torch.int64, device=self.device)
        cen = torch.empty((nv, 3), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_lattice_export(self._handle, ptr(addr), ptr(cen), stream_ptr(self.device)))
        return addr, cen

    def radius_features(self, query_cloud, radii, out_dtype=np.float64, descriptors="reference", algorithm=0):
        """(Nq, C*len(radii)) features for several radii sharing this lattice's edge; CUDA tensor."""
        self._sync_streams()
        validate_cloud(query_cloud, "query_cloud")
        np_out, out_code = _out_code(out_dtype)
        ncol, mask = _ncol(descriptors)
        q, qc = device_cloud(query_cloud, self.device)
        radii_arr, radii_p = _lib.f64_array(list(radii))
        out = torch.zeros((q.shape[0], ncol * len(radii_arr)), dtype=_TORCH_OUT[np_out], device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_radius_features(
                self._handle, ptr(q), qc, q.shape[0], radii_p, len(radii_arr), ptr(out), out_code, out.shape[1], 0,
                mask, int(algorithm), stream_ptr(self.device)))
        return out

    def radius_sets(self, query_cloud, radius):
        """CSR neighbor index sets (offsets int64 (Nq+1,), indices int32) -- what query_ball_tree returns."""
        self._sync_streams()
        validate_cloud(query_cloud, "query_cloud")
        q, qc = device_cloud(query_cloud, self.device)
        nq = q.shape[0]
        offsets = torch.zeros(nq + 1, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            s = stream_ptr(self.device)
            _lib.check(_lib.lib().nbr_radius_sets(self._handle, ptr(q), qc, nq, float(radius), ptr(offsets), None, s))
            total = int(offsets[-1].item())
            indices = torch.empty(max(total, 1), dtype=torch.int32, device=self.device)
            _lib.check(_lib.lib().nbr_radius_sets(self._handle, ptr(q), qc, nq, float(radius), ptr(offsets),
                                                  ptr(indices), s))
        return offsets, indices[:total]

    def knn(self, query_cloud, k, ks=None, out_dtype=np.float64, descriptors="reference"):
        """
        k nearest voxels per query in (squared distance, index) order.
        returns (indices int32 (Nq,k), d2 float64 (Nq,k)[, features (Nq, C*len(ks))]).
        """
        self._sync_streams()
        validate_cloud(query_cloud, "query_cloud")
        q, qc = device_cloud(query_cloud, self.device)
        nq = q.shape[0]
        idx = torch.empty((nq, k), dtype=torch.int32, device=self.device)
        d2 = torch.empty((nq, k), dtype=torch.float64, device=self.device)
        feats = None
        np_out, out_code = _out_code(out_dtype)
        ncol, mask = _ncol(descriptors)
        ks_arr = None
        if ks is not None:
            ks_arr = np.ascontiguousarray(ks, dtype=np.int32)
            if np.any(np.diff(ks_arr) <= 0) or ks_arr[-1] != k:
                raise ValueError("ks must be ascending and end at k")
            feats = torch.zeros((nq, ncol * len(ks_arr)), dtype=_TORCH_OUT[np_out], device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().nbr_knn(
                self._handle, ptr(q), qc, nq, int(k), ptr(idx), ptr(d2),
                ks_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if ks_arr is not None else None,
                len(ks_arr) if ks_arr is not None else 0, ptr(feats), out_code,
                feats.shape[1] if feats is not None else 0, 0, mask, stream_ptr(self.device)))
        return (idx, d2) if feats is None else (idx, d2, feats)


def knn_points(query_cloud, search_cloud, k, ks=None, out_dtype=np.float64, descriptors="reference", cell_edge=0.0):
    """
    k nearest RAW points of the search cloud per query (no voxel filter; the legacy sspedge = 0), in
    (squared distance, index) order.  returns (indices int32 (Nq,k) into search_cloud, d2 float64 (Nq,k)
    [, features (Nq, C*len(ks))]) as CUDA tensors.
    """
    validate_cloud(query_cloud, "query_cloud")
    validate_cloud(search_cloud, "search_cloud")
    if search_cloud.shape[0] < 1:
        raise ValueError("need at least 1 search point")
    require_cuda()
    q, qc = device_cloud(query_cloud)
    s, sc = (q, qc) if search_cloud is query_cloud else device_cloud(search_cloud, q.device)
    nq = q.shape[0]
    idx = torch.empty((nq, k), dtype=torch.int32, device=q.device)
    d2 = torch.empty((nq, k), dtype=torch.float64, device=q.device)
    np_out, out_code = _out_code(out_dtype)
    ncol, mask = _ncol(descriptors)
    feats, ks_arr = None, None
    if ks is not None:
        ks_arr = np.ascontiguousarray(ks, dtype=np.int32)
        if np.any(np.diff(ks_arr) <= 0) or ks_arr[-1] != k:
            raise ValueError("ks must be ascending and end at k")
        feats = torch.zeros((nq, ncol * len(ks_arr)), dtype=_TORCH_OUT[np_out], device=q.device)
    with torch.cuda.device(q.device):
        _lib.check(_lib.lib().nbr_knn_points(
            ptr(s), sc, s.shape[0], ptr(q), qc, nq, int(k), float(cell_edge), ptr(idx), ptr(d2),
            ks_arr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if ks_arr is not None else None,
            len(ks_arr) if ks_arr is not None else 0, ptr(feats), out_code, feats.shape[1] if feats is not None else 0, 0, mask,
            stream_ptr(q.device)))
    return (idx, d2) if feats is None else (idx, d2, feats)


def knn_features(query_cloud, search_cloud, edge_length, ks, out_dtype=np.float64, descriptors="reference"):
    """
    multiscale kNN features (extension; BASELINE config 3): the search cloud is voxel-filtered at
    edge_length, and for each k in ks the reference's 4 columns are computed over the k nearest
    voxels, ties broken by (distance, index).  (Nq, C*len(ks)); numpy in -> numpy out.
    edge_length = 0 searches the raw points instead (the legacy convention sspedge = 0, prototypes/mso.py:277).
    """
    ks = sorted(int(k) for k in ks)
    if not edge_length:
        _, _, feats = knn_points(query_cloud, search_cloud, ks[-1], ks=ks, out_dtype=out_dtype, descriptors=descriptors)
        if is_torch(query_cloud):
            return feats if query_cloud.is_cuda else feats.cpu()
        return feats.cpu().numpy()
    index = LatticeIndex(search_cloud, edge_length, indexed=True)
    try:
        _, _, feats = index.knn(query_cloud, ks[-1], ks=ks, out_dtype=out_dtype, descriptors=descriptors)
    finally:
        index.close()
    if is_torch(query_cloud):
        return feats if query_cloud.is_cuda else feats.cpu()
    return feats.cpu().numpy()


def vector_field_features(query_cloud, search_cloud, search_vectors, edge_length, radii, out_dtype=np.float32):
    """
    vector-field multiscale operator (extension; legacy precedent V_MSO, nimrud/prototypes/mso.py:12-257):
    every search point carries a feature vector (search_vectors, (Ns, F)); the vectors are averaged per voxel
    of the search cloud's lattice at `edge_length`, and for every query and every radius the mean of the voxel
    vectors over the voxels within the radius (inclusive, as the eigenfeature path) is returned:
    (Nq, F * len(radii)), scale-major; zeros where a neighborhood is empty.  numpy in -> numpy out.
    """
    validate_cloud(query_cloud, "query_cloud")
    if search_cloud.ndim != 2 or search_cloud.shape[1] != 3:
        raise ValueError("only 3D search clouds can be indexed")
    np_out, out_code = _out_code(out_dtype)
    index = LatticeIndex(search_cloud, edge_length, indexed=True)
    try:
        dev = index.device
        vec = search_vectors if is_torch(search_vectors) else torch.from_numpy(np.ascontiguousarray(search_vectors))
        vec = vec.to(device=dev, dtype=torch.float32).reshape(search_cloud.shape[0], -1).contiguous()
        F = int(vec.shape[1])
        if F < 1:
            raise ValueError("search_vectors needs at least one component")
        q, qc = device_cloud(query_cloud, dev)
        radii_l = [float(r) for r in radii]
        voxvec = torch.zeros((max(index.n_voxels, 1), F), dtype=torch.float32, device=dev)
        out = torch.zeros((q.shape[0], F * len(radii_l)), dtype=_TORCH_OUT[np_out], device=dev)
        with torch.cuda.device(dev):
            s = stream_ptr(dev)
            _lib.check(_lib.lib().nbr_voxel_vector_means(index._handle, ptr(index._search), index._code,
                                                         index._search.shape[0], ptr(vec), F, ptr(voxvec), s))
            for k, r in enumerate(radii_l):
                _lib.check(_lib.lib().nbr_radius_vector_means(index._handle, ptr(q), qc, q.shape[0], r, ptr(voxvec), F,
                                                              ptr(out), out_code, out.shape[1], k * F, s))
    finally:
        index.close()
    if is_torch(query_cloud):
        return out if query_cloud.is_cuda else out.cpu()
    return out.cpu().numpy()
