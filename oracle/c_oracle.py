"""
ctypes binding of oracle/oracle.c (plain-C restatement; TEST INFRASTRUCTURE ONLY -- see the header
of oracle.c for what may load it).  build with `make -C oracle` or __graft_entry__.build().
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")

_lib = None
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force=False):
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-B", "_build/liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def grid_widths(points, edge):
    pts = _c(points, np.float64)
    minc = np.zeros(3); maxc = np.zeros(3); widths = np.zeros(3, dtype=np.int64)
    rc = lib().orc_grid_widths(_p(pts, _f64p), ctypes.c_int64(len(pts)), ctypes.c_double(edge),
                               _p(minc, _f64p), _p(maxc, _f64p), _p(widths, _i64p))
    if rc == 1:
        raise ValueError("need at least 2 points to define a voxel grid")
    if rc == 2:
        raise ValueError("edge length is too small to address this space")
    return minc, maxc, widths


def unique_voxels(points, minc, edge, widths):
    pts = _c(points, np.float64)
    keys = np.zeros(len(pts), dtype=np.int64)
    centres = np.zeros((len(pts), 3))
    f = lib().orc_unique_voxels
    f.restype = ctypes.c_int64
    nv = f(_p(pts, _f64p), ctypes.c_int64(len(pts)), _p(_c(minc, np.float64), _f64p),
           ctypes.c_double(edge), _p(_c(widths, np.int64), _i64p), _p(keys, _i64p), _p(centres, _f64p))
    return keys[:nv].copy(), centres[:nv].copy()


def radius_features(query, ukeys, minc, edge, widths, radius):
    q = _c(query, np.float64)
    out = np.zeros((len(q), 4))
    rc = lib().orc_radius_features(_p(q, _f64p), ctypes.c_int64(len(q)), _p(_c(ukeys, np.int64), _i64p),
                                   ctypes.c_int64(len(ukeys)), _p(_c(minc, np.float64), _f64p),
                                   ctypes.c_double(edge), _p(_c(widths, np.int64), _i64p),
                                   ctypes.c_double(radius), _p(out, _f64p))
    assert rc == 0
    return out


def radius_sets(query, ukeys, minc, edge, widths, radius):
    q = _c(query, np.float64)
    uk = _c(ukeys, np.int64); mc = _c(minc, np.float64); wd = _c(widths, np.int64)
    offsets = np.zeros(len(q) + 1, dtype=np.int64)
    args = (_p(q, _f64p), ctypes.c_int64(len(q)), _p(uk, _i64p), ctypes.c_int64(len(uk)),
            _p(mc, _f64p), ctypes.c_double(edge), _p(wd, _i64p), ctypes.c_double(radius))
    assert lib().orc_radius_sets(*args, _p(offsets, _i64p), None) == 0
    indices = np.zeros(max(int(offsets[-1]), 1), dtype=np.int64)
    assert lib().orc_radius_sets(*args, _p(offsets, _i64p), _p(indices, _i64p)) == 0
    return offsets, indices[:offsets[-1]]


def process(query, search, edges, radii, threads=None):
    if threads is not None:
        lib().orc_set_threads(int(threads))
    q = _c(query, np.float64); s = _c(search, np.float64)
    e = _c(edges, np.float64); r = _c(radii, np.float64)
    assert len(e) == len(r), "edge_lengths and radii should be equal-length sequences."
    out = np.zeros((len(q), 4 * len(e)))
    rc = lib().orc_process(_p(q, _f64p), ctypes.c_int64(len(q)), _p(s, _f64p), ctypes.c_int64(len(s)),
                           _p(e, _f64p), _p(r, _f64p), ctypes.c_int32(len(e)), _p(out, _f64p))
    if rc == 1:
        raise ValueError("need at least 2 points to define a voxel grid")
    if rc == 2:
        raise ValueError("edge length is too small to address this space")
    return out


def knn(query, points, k):
    q = _c(query, np.float64); p = _c(points, np.float64)
    idx = np.zeros((len(q), k), dtype=np.int64)
    d2 = np.zeros((len(q), k))
    assert lib().orc_knn(_p(q, _f64p), ctypes.c_int64(len(q)), _p(p, _f64p), ctypes.c_int64(len(p)),
                         ctypes.c_int32(k), _p(idx, _i64p), _p(d2, _f64p)) == 0
    return idx, d2


def knn_features(query, points, knn_idx, ks):
    q = _c(query, np.float64); p = _c(points, np.float64)
    idx = _c(knn_idx, np.int64); ks = _c(ks, np.int32)
    out = np.zeros((len(q), 4 * len(ks)))
    assert lib().orc_knn_features(_p(q, _f64p), ctypes.c_int64(len(q)), _p(p, _f64p), _p(idx, _i64p),
                                  ctypes.c_int32(idx.shape[1]), _p(ks, _i32p), ctypes.c_int32(len(ks)),
                                  _p(out, _f64p)) == 0
    return out


def num_threads():
    return int(lib().orc_num_threads())
