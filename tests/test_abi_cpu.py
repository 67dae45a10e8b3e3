"""
CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, its host-only entry points give the reference's known answers, and the product refuses
to run (instead of falling back) without a GPU.
"""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from nimrud_b200 import _lib


def header_symbols():
    text = open(os.path.join(ROOT, "include", "nimrud_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = header_symbols()
    assert len(names) >= 20
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), "libnimrud_b200.so does not export %s" % name
    # the ctypes table covers the whole header, and nothing else
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_string():
    lib = _lib.lib()
    assert lib.nbr_version() >= 100
    assert isinstance(_lib.last_error(), str)


def grid(lo, hi, edge, ndim=3):
    from nimrud_b200.geometry import grid_from_bbox
    return grid_from_bbox(lo, hi, edge, ndim)


def test_grid_known_answers_host_only():
    # nimrud/utils/tests/geometry_tests.py:84-158
    for dim in (2, 3):
        g = grid([0, 0, 0][:dim], [100, 100, 100][:dim], 0.001, dim)
        assert list(g.widths[:dim]) == [17, 17, 17][:dim]
        assert list(g.shifts[:dim]) == [0, 17, 34][:dim]
        with pytest.raises(ValueError):
            grid([0, 0, 0][:dim], [100, 100, 100][:dim], 0.00001 if dim == 3 else 0.00000001, dim)
        g = grid([0, 0, 0][:dim], [100, 100, 100][:dim], 1, dim)
        assert list(g.widths[:dim]) == [7, 7, 7][:dim]
        assert list(g.min_corner[:dim]) == [-0.5] * dim and list(g.max_corner[:dim]) == [100.5] * dim


def test_grid_matches_oracle_on_random_boxes():
    from oracle import nimrud_oracle as O
    rs = np.random.RandomState(0)
    for _ in range(200):
        pts = rs.rand(2, 3) * rs.choice([1.0, 10.0, 1000.0]) + rs.randn(3) * 5
        edge = float(rs.choice([0.01, 0.1, 0.25, 1.0, 3.0]))
        p = O.grid_params(pts, edge)
        g = grid(pts.min(0), pts.max(0), edge)
        assert list(g.widths) == list(p.widths)
        assert np.array_equal(np.array(g.min_corner[:]), p.minimum_corner)
        assert np.array_equal(np.array(g.max_corner[:]), p.maximum_corner)


def test_argument_validation_needs_no_gpu():
    from nimrud_b200 import multiscale
    pts = np.random.rand(10, 3)
    with pytest.raises(AssertionError):
        multiscale.process_single_core(pts, pts, [0.1, 0.2], [0.5])
    with pytest.raises(ValueError):
        multiscale.process_single_core(pts, np.random.rand(1, 3), [0.1], [0.5])
    with pytest.raises(ValueError):
        multiscale.process_single_core(pts, np.random.rand(10, 4), [0.1], [0.5])
    with pytest.raises(ValueError):
        multiscale.process_single_core(pts, np.random.rand(10), [0.1], [0.5])
    with pytest.raises(ValueError):
        multiscale.process_single_core(np.random.rand(10, 4), pts, [0.1], [0.5])


def test_gather_staging_layout_host_only():
    """staging buffer of the feature all-gather: rows, then (256-byte aligned) one uint32 row number per row"""
    lib = _lib.lib()
    assert lib.nbr_gather_staging_bytes(0, 80) == 0
    assert lib.nbr_gather_staging_bytes(1, 80) == 256 + 4
    assert lib.nbr_gather_staging_bytes(16, 80) == 1280 + 64                 # 1280 is a multiple of 256 already
    assert lib.nbr_gather_staging_bytes(80_000_000, 80) == 6_400_000_000 + 320_000_000
    assert lib.nbr_gather_staging_bytes(-1, 80) == 0
    # null / bad arguments fail before any CUDA call
    assert lib.nbr_tile_step_gather(None, None, _lib.F32, 0, None, None, 0, None, 0, _lib.F32, 0, None, None, None, None) == _lib.ERR_INVALID
    assert lib.nbr_gather_finish(None, None) == _lib.ERR_INVALID
    assert lib.nbr_gather_unpermute(None, None, 80, None, None) == _lib.ERR_INVALID


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from nimrud_b200 import geometry, multiscale
    pts = np.random.rand(10, 3)
    with pytest.raises(RuntimeError):
        multiscale.process_single_core(pts, pts, [0.1], [0.5])
    with pytest.raises(RuntimeError):
        geometry.VoxelFilter(pts, 0.5)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "nimrud_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower(), "%s mentions the oracle" % os.path.join(dirpath, f)
