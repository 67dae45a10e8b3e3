"""
GPU-backed mirror of `nimrud.utils.geometry.VoxelFilter` (reference: nimrud/utils/geometry.py:16-154).

Same constructor, attributes and methods; the arithmetic runs in the CUDA library
(nbr_bbox, nbr_grid_from_bbox, nbr_voxel_addresses, nbr_sort_u64 + nbr_unique_u64,
nbr_voxel_centres).  numpy in -> numpy out, CUDA tensors in -> CUDA tensors out.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._util import device_cloud, is_torch, ptr, require_cuda, stream_ptr

MAX_ADDRESS_LENGTH = 64


def grid_from_bbox(lo, hi, edge_length, ndim=3):
    """nbr_grid (ctypes struct) for a bounding box; ValueError if it needs more than 64 address bits."""
    lo3 = np.zeros(3); hi3 = np.zeros(3)
    lo3[:ndim] = np.asarray(lo, dtype=np.float64)[:ndim]
    hi3[:ndim] = np.asarray(hi, dtype=np.float64)[:ndim]
    grid = _lib.Grid()
    f64p = ctypes.POINTER(ctypes.c_double)
    _lib.check(_lib.lib().nbr_grid_from_bbox(lo3.ctypes.data_as(f64p), hi3.ctypes.data_as(f64p),
                                             float(edge_length), int(ndim), ctypes.byref(grid)))
    return grid


def cloud_bbox(points_dev, dtype_code):
    """(lo, hi) float64 numpy arrays of a CUDA (n, ndim) cloud."""
    n, ndim = points_dev.shape
    out = torch.empty(6, dtype=torch.float64, device=points_dev.device)
    with torch.cuda.device(points_dev.device):     # every library call runs with the data's device current
        _lib.check(_lib.lib().nbr_bbox(ptr(points_dev), dtype_code, n, ndim, ptr(out), stream_ptr(points_dev.device)))
    box = out.cpu().numpy()
    return box[:ndim].copy(), box[3:3 + ndim].copy()


class VoxelFilter(object):
    """
    given a 2d or 3d point cloud, a cubic grid of the given edge length enclosing it; converts
    coordinates to packed 64-bit cell addresses and back.  (reference: utils/geometry.py:16-21)
    """

    def __init__(self, points, edge_length):
        # utils/geometry.py:30-35
        if points.ndim != 2:
            raise ValueError("wrong point cloud array shape")
        elif points.shape[1] not in [2, 3]:
            raise ValueError("only 2D and 3D spaces supported")
        elif points.shape[0] < 2:
            raise ValueError("need at least 2 points to define a voxel grid")
        require_cuda()
        dev, code = device_cloud(points)
        self._ndim = int(points.shape[1])
        self._device = dev.device
        lo, hi = cloud_bbox(dev, code)
        self.edge_length = edge_length
        self._grid = grid_from_bbox(lo, hi, edge_length, self._ndim)       # ValueError on > 64 bits
        self.minimum_corner = np.array(self._grid.min_corner[:self._ndim])  # utils/geometry.py:37
        self.maximum_corner = np.array(self._grid.max_corner[:self._ndim])  # utils/geometry.py:38
        self.widths = np.array(self._grid.widths[:self._ndim], dtype=np.int64)
        self.shifts = np.array(self._grid.shifts[1:self._ndim], dtype=np.int64)
        # utils/geometry.py:74-77
        self.masks = [((1 << int(w)) - 1) << int(s) for w, s in zip(self.widths, self._grid.shifts[:self._ndim])]

    @classmethod
    def from_bbox(cls, lo, hi, edge_length, device=None):
        """filter anchored on an explicit bounding box (multi-GPU: the all-reduced box)."""
        require_cuda()
        self = cls.__new__(cls)
        self._ndim = len(lo)
        self._device = torch.device(device or "cuda")
        self.edge_length = edge_length
        self._grid = grid_from_bbox(lo, hi, edge_length, self._ndim)
        self.minimum_corner = np.array(self._grid.min_corner[:self._ndim])
        self.maximum_corner = np.array(self._grid.max_corner[:self._ndim])
        self.widths = np.array(self._grid.widths[:self._ndim], dtype=np.int64)
        self.shifts = np.array(self._grid.shifts[1:self._ndim], dtype=np.int64)
        self.masks = [((1 << int(w)) - 1) << int(s) for w, s in zip(self.widths, self._grid.shifts[:self._ndim])]
        return self

    # ------------------------------------------------------------------
    def _check_shape(self, points):
        """the shape half of utils/geometry.py:83-99; the bounds half runs on the GPU."""
        pts = points if is_torch(points) else np.asarray(points)
        if pts.ndim == 1:
            pts = pts.reshape(1, -1)
        if pts.ndim != 2:
            raise ValueError("wrong array shape")
        if pts.shape[1] != self._ndim:
            raise ValueError("wrong number of spatial dimensions")
        return pts

    def _check_in_bounds(self, points):
        pts = self._check_shape(points)
        self._addresses_dev(pts)      # raises ValueError when a point is outside
        return pts

    def _addresses_dev(self, pts):
        dev, code = device_cloud(pts, self._device)
        n = dev.shape[0]
        addr = torch.empty(n, dtype=torch.int64, device=dev.device)
        oob = torch.zeros(1, dtype=torch.int32, device=dev.device)
        with torch.cuda.device(dev.device):
            _lib.check(_lib.lib().nbr_voxel_addresses(ptr(dev), code, n, ctypes.byref(self._grid), ptr(addr), ptr(oob),
                                                      stream_ptr(dev.device)))
        if int(oob.item()) != 0:
            raise ValueError("some points fall outside filter bounding region")   # utils/geometry.py:96-97
        return addr

    def coordinate_to_address(self, points):
        """real-world coordinates -> packed integer addresses (utils/geometry.py:103-116)."""
        pts = self._check_shape(points)
        addr = self._addresses_dev(pts)
        return addr if is_torch(points) else addr.cpu().numpy()

    def _centres_dev(self, addr_dev):
        n = addr_dev.shape[0]
        out = torch.empty((n, self._ndim), dtype=torch.float64, device=addr_dev.device)
        with torch.cuda.device(addr_dev.device):
            _lib.check(_lib.lib().nbr_voxel_centres(ptr(addr_dev), n, ctypes.byref(self._grid), ptr(out),
                                                    stream_ptr(addr_dev.device)))
        return out

    def address_to_coordinate(self, addresses):
        """packed addresses -> voxel centre coordinates (utils/geometry.py:120-138)."""
        if is_torch(addresses):
            a = addresses.to(self._device, torch.int64).reshape(-1).contiguous()
            return self._centres_dev(a)
        a = torch.from_numpy(np.atleast_1d(np.asarray(addresses, dtype=np.int64)).copy()).to(self._device)
        return self._centres_dev(a).cpu().numpy()

    def _unique_addresses_dev(self, pts):
        addr = self._addresses_dev(pts)
        n = addr.shape[0]
        tmp = torch.empty_like(addr)
        bits = int(self.widths.sum())
        s = stream_ptr(addr.device)
        count = torch.zeros(1, dtype=torch.int64, device=addr.device)
        with torch.cuda.device(addr.device):
            _lib.check(_lib.lib().nbr_sort_u64(ptr(addr), ptr(tmp), n, 0, bits, s))
            _lib.check(_lib.lib().nbr_unique_u64(ptr(addr), n, ptr(tmp), ptr(count), s))
        return tmp[:int(count.item())]

    def unique_addresses(self, points):
        """sorted unique addresses of the cells that hold a point (np.unique at utils/geometry.py:150)."""
        pts = self._check_shape(points)
        uniq = self._unique_addresses_dev(pts)
        return uniq.clone() if is_torch(points) else uniq.cpu().numpy()

    def unique_voxels(self, points):
        """unique centre coordinates of all grid cells that contain a point (utils/geometry.py:142-154)."""
        pts = self._check_shape(points)
        centres = self._centres_dev(self._unique_addresses_dev(pts).contiguous())
        return centres if is_torch(points) else centres.cpu().numpy()

    def find_neighbors(self, address):
        raise NameError("find_neighbors not implemented yet")          # utils/geometry.py:158-164

    def find_facing_neighbors(self, address):
        raise NameError("find_facing_neighbors not implemented yet")   # utils/geometry.py:166-172
