"""host-side helpers shared by the shims: moving clouds to the device, stream handles."""
import ctypes

import numpy as np
import torch

from . import _lib


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("nimrud_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")


def is_torch(x):
    return isinstance(x, torch.Tensor)


def validate_cloud(points, what, allow_2d=False):
    """shape checks with the reference's exception type (utils/geometry.py:30-35)."""
    if points.ndim != 2:
        raise ValueError("wrong point cloud array shape")
    dims = (2, 3) if allow_2d else (3,)
    if points.shape[1] not in dims:
        raise ValueError("only 2D and 3D spaces supported" if allow_2d
                         else "%s must have shape (N, 3)" % what)


def device_cloud(points, device=None):
    """-> (contiguous CUDA tensor of float32 or float64, dtype code).  other dtypes become float64."""
    require_cuda()
    if is_torch(points):
        t = points
        if t.dtype not in (torch.float32, torch.float64):
            t = t.to(torch.float64)
        if not t.is_cuda:
            t = t.to(device or "cuda")
    else:
        arr = np.asarray(points)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(arr)).to(device or "cuda")
    t = t.contiguous()
    return t, (_lib.F32 if t.dtype == torch.float32 else _lib.F64)


def host_cloud(points):
    """-> (C-contiguous numpy float32/float64 array, dtype code)"""
    arr = np.asarray(points)
    if arr.dtype not in (np.float32, np.float64):
        arr = arr.astype(np.float64)
    arr = np.ascontiguousarray(arr)
    return arr, (_lib.F32 if arr.dtype == np.float32 else _lib.F64)


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None
