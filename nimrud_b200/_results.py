"""
result arrays of the host-buffer path.

a fresh pageable (N, 4*S) float64 array costs more in page faults (the kernel zeroes 1.6 GB for 10M points x 5
scales) than its rows cost on the PCIe link.  results are therefore numpy arrays over buffers from the C library
(2 MB-aligned, huge pages requested), and a buffer whose array (and every view of it) has been garbage collected is
kept for the next call of the same size: a recycled buffer is already mapped.  NBR_RESULT_POOL_MB bounds what is kept (default 4096; 0 disables the pool: plain np.empty).
"""
import ctypes
import os
import threading
import weakref

import numpy as np

from . import _lib

_lock = threading.Lock()
_free = {}          # nbytes -> [pointer, ...]
_kept = 0


def _limit():
    return int(float(os.environ.get("NBR_RESULT_POOL_MB", "4096")) * (1 << 20))


def _release(ptr, nbytes):
    global _kept
    with _lock:
        if _kept + nbytes <= _limit():
            _free.setdefault(nbytes, []).append(ptr)
            _kept += nbytes
            return
    try:
        _lib.lib().nbr_host_free(ctypes.c_void_p(ptr))
    except Exception:
        pass


def empty(shape, dtype):
    """uninitialised array of `shape`: recycled when the pool is enabled, np.empty otherwise."""
    global _kept
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes < (1 << 20) or nbytes > _limit():
        return np.empty(shape, dtype=dtype)
    ptr = None
    with _lock:
        bucket = _free.get(nbytes)
        if bucket:
            ptr = bucket.pop()
            _kept -= nbytes
    if ptr is None:
        out = ctypes.c_void_p()
        if _lib.lib().nbr_host_alloc(nbytes, 0, ctypes.byref(out)) != _lib.OK or not out.value:
            return np.empty(shape, dtype=dtype)
        ptr = out.value
    raw = (ctypes.c_char * nbytes).from_address(ptr)
    weakref.finalize(raw, _release, ptr, nbytes)          # runs when the array and all its views are gone
    return np.frombuffer(raw, dtype=dtype).reshape(shape)


def trim():
    """free every cached buffer."""
    global _kept
    with _lock:
        ptrs = [p for bucket in _free.values() for p in bucket]
        _free.clear()
        _kept = 0
    for p in ptrs:
        _lib.lib().nbr_host_free(ctypes.c_void_p(p))
