"""
ctypes binding of include/nimrud_b200.h.  Loads nimrud_b200/lib/libnimrud_b200.so (built in-tree by
__graft_entry__.build() or nimrud_b200/csrc/build.sh) and fails loudly when it is missing.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NIMRUD_B200_LIB") or os.path.join(HERE, "lib", "libnimrud_b200.so")   # env: A/B builds

OK = 0
ERR_INVALID, ERR_TOO_FEW_POINTS, ERR_ADDRESS_BITS, ERR_CUDA, ERR_UNSUPPORTED, ERR_OUT_OF_BOUNDS, ERR_CAPACITY = 1, 2, 3, 4, 5, 6, 7
F32, F64 = 0, 1
DESC_REFERENCE, DESC_EXTENDED = 0, 1
LATTICE_INDEXED = 1

c_i64, c_i32, c_f64, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_double, ctypes.c_void_p


class Grid(ctypes.Structure):
    """struct nbr_grid"""
    _fields_ = [("min_corner", c_f64 * 3), ("max_corner", c_f64 * 3), ("edge", c_f64),
                ("widths", c_i32 * 3), ("shifts", c_i32 * 3), ("ndim", c_i32), ("reserved", c_i32)]


# every symbol include/nimrud_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "nbr_last_error": (ctypes.c_char_p, []),
    "nbr_version": (ctypes.c_int, []),
    "nbr_trim_memory": (ctypes.c_int, []),
    "nbr_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.c_int, ctypes.POINTER(c_vp)]),
    "nbr_host_free": (ctypes.c_int, [c_vp]),
    "nbr_kernel_launches": (c_i64, []),
    "nbr_debug_bounds_violations": (c_i64, []),
    "nbr_timing_enable": (None, [ctypes.c_int]),
    "nbr_timing_read": (ctypes.c_int, [ctypes.POINTER(c_f64)]),
    "nbr_bbox": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.c_int, c_vp, c_vp]),
    "nbr_grid_from_bbox": (ctypes.c_int, [ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_f64, ctypes.c_int,
                                          ctypes.POINTER(Grid)]),
    "nbr_voxel_addresses": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.POINTER(Grid), c_vp, c_vp, c_vp]),
    "nbr_sort_u64": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_vp]),
    "nbr_sort_pairs_u64_u32": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, ctypes.c_int, ctypes.c_int, c_vp]),
    "nbr_unique_u64": (ctypes.c_int, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "nbr_voxel_centres": (ctypes.c_int, [c_vp, c_i64, ctypes.POINTER(Grid), c_vp, c_vp]),
    "nbr_exclusive_scan_u32": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "nbr_exclusive_scan_i64": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "nbr_lattice_create": (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp, ctypes.c_int, c_i64, ctypes.POINTER(Grid),
                                          ctypes.c_int, c_vp]),
    "nbr_lattice_destroy": (None, [c_vp]),
    "nbr_lattice_info": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), ctypes.POINTER(c_i64)]),
    "nbr_lattice_export": (ctypes.c_int, [c_vp, c_vp, c_vp, c_vp]),
    "nbr_radius_features": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, ctypes.POINTER(c_f64), c_i32, c_vp,
                                           ctypes.c_int, c_i64, c_i32, c_i32, c_i32, c_vp]),
    "nbr_radius_sets": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_f64, c_vp, c_vp, c_vp]),
    "nbr_knn": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_i32, c_vp, c_vp, ctypes.POINTER(c_i32), c_i32,
                               c_vp, ctypes.c_int, c_i64, c_i32, c_i32, c_vp]),
    "nbr_knn_points": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, ctypes.c_int, c_i64, c_i32, c_f64, c_vp, c_vp,
                                      ctypes.POINTER(c_i32), c_i32, c_vp, ctypes.c_int, c_i64, c_i32, c_i32, c_vp]),
    "nbr_voxel_vector_means": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_vp, c_i32, c_vp, c_vp]),
    "nbr_radius_vector_means": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_f64, c_vp, c_i32, c_vp, ctypes.c_int,
                                               c_i64, c_i32, c_vp]),
    "nbr_halo_count": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.POINTER(c_f64), c_i32, c_vp, c_vp]),
    "nbr_halo_fill": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.POINTER(c_f64), c_i32, ctypes.POINTER(c_i64),
                                     c_vp, c_vp, c_vp]),
    "nbr_brick_origin": (ctypes.c_int, [ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_f64, ctypes.POINTER(c_f64)]),
    "nbr_order_cloud": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_f64,
                                       c_vp, c_vp, c_vp]),
    "nbr_multiscale_features_tile": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_vp, c_i64, ctypes.POINTER(c_f64),
                                                    ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), ctypes.POINTER(c_f64),
                                                    c_i32, c_vp, ctypes.c_int, c_i32, ctypes.POINTER(c_i64), c_vp]),
    "nbr_mailbox_create": (ctypes.c_int, [ctypes.POINTER(c_vp), c_i32, c_i32, ctypes.c_int, c_i64]),
    "nbr_mailbox_destroy": (None, [c_vp]),
    "nbr_mailbox_disconnect": (ctypes.c_int, [c_vp, c_i32]),
    "nbr_mailbox_ipc_handle": (ctypes.c_int, [c_vp, c_vp]),
    "nbr_mailbox_connect_ipc": (ctypes.c_int, [c_vp, c_i32, c_vp]),
    "nbr_mailbox_connect_local": (ctypes.c_int, [c_vp, c_i32, c_vp]),
    "nbr_mailbox_set_peer_capacity": (ctypes.c_int, [c_vp, c_i32, c_i64]),
    "nbr_tile_box_publish": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_vp]),
    "nbr_tile_boxes_wait": (ctypes.c_int, [c_vp, ctypes.POINTER(c_f64), c_vp]),
    "nbr_halo_push": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, ctypes.POINTER(c_f64), c_f64, c_vp]),
    "nbr_halo_wait": (ctypes.c_int, [c_vp, c_vp]),
    "nbr_mailbox_rows": (c_vp, [c_vp]),
    "nbr_mailbox_count_dev": (c_vp, [c_vp]),
    "nbr_mailbox_read": (ctypes.c_int, [c_vp, c_vp, c_i64, ctypes.POINTER(c_i64), c_vp]),
    "nbr_mailbox_status": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "nbr_multiscale_features_tile_mb": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_vp, ctypes.POINTER(c_f64),
                                                       ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), ctypes.POINTER(c_f64),
                                                       c_i32, c_vp, ctypes.c_int, c_i32, ctypes.POINTER(c_i64), c_vp]),
    "nbr_tile_step": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_i32, c_vp,
                                     ctypes.c_int, c_i32, ctypes.POINTER(c_f64), ctypes.POINTER(c_i64), c_vp]),
    "nbr_mailbox_gather_alloc": (ctypes.c_int, [c_vp, ctypes.c_uint64]),
    "nbr_mailbox_gather_ipc_handle": (ctypes.c_int, [c_vp, c_vp]),
    "nbr_mailbox_gather_connect_ipc": (ctypes.c_int, [c_vp, c_i32, c_vp, ctypes.c_uint64]),
    "nbr_mailbox_gather_connect_local": (ctypes.c_int, [c_vp, c_i32, c_vp]),
    "nbr_mailbox_gather_ptr": (c_vp, [c_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "nbr_multiscale_features_tile_mb_gather": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, c_vp, ctypes.POINTER(c_f64),
                                                              ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), ctypes.POINTER(c_f64),
                                                              c_i32, c_vp, ctypes.c_int, c_i32, c_i64, c_i64, ctypes.POINTER(c_i64), c_vp]),
    "nbr_gather_finish": (ctypes.c_int, [c_vp, c_vp]),
    "nbr_gather_unpermute": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64), c_i64, c_vp, c_vp]),
    "nbr_gather_staging_bytes": (ctypes.c_uint64, [c_i64, c_i64]),
    "nbr_tile_step_gather": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_i32, c_vp,
                                            c_i64, ctypes.c_int, c_i32, ctypes.POINTER(c_f64), ctypes.POINTER(c_i64), ctypes.POINTER(c_i64),
                                            c_vp]),
    "nbr_tile_step_host": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int, c_i64, ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_i32, c_vp,
                                          ctypes.c_int, c_i32, ctypes.POINTER(c_f64)]),
    "nbr_multiscale_features": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, ctypes.c_int, c_i64,
                                               ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_i32, c_vp,
                                               ctypes.c_int, c_i32, ctypes.POINTER(c_f64), ctypes.POINTER(c_i64),
                                               c_vp]),
    "nbr_multiscale_features_host": (ctypes.c_int, [c_vp, ctypes.c_int, c_i64, c_vp, ctypes.c_int, c_i64,
                                                    ctypes.POINTER(c_f64), ctypes.POINTER(c_f64), c_i32, c_vp,
                                                    ctypes.c_int, c_i32, ctypes.POINTER(c_i64)]),
}

_lib = None


class LibraryMissing(RuntimeError):
    pass


def lib():
    """the loaded C-ABI library.  raises LibraryMissing (never falls back) if it was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise LibraryMissing(
                "nimrud_b200: %s not found. build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or nimrud_b200/csrc/build.sh. there is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)     # AttributeError if the header and the library disagree
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error():
    return lib().nbr_last_error().decode("utf-8", "replace")


def check(rc):
    """map C status codes onto the exception types the reference raises."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (ERR_TOO_FEW_POINTS, ERR_ADDRESS_BITS, ERR_OUT_OF_BOUNDS, ERR_INVALID):
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == ERR_CAPACITY:
        raise RuntimeError(msg)
    raise RuntimeError("nimrud_b200 CUDA error: " + msg)


def f64_array(values):
    arr = np.ascontiguousarray(values, dtype=np.float64)
    return arr, arr.ctypes.data_as(ctypes.POINTER(c_f64))
