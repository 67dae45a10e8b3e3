import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, "golden_%s.npz" % name))


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as mod
    mod.build()
    return mod


# tolerances of the parity bar (BASELINE.json north_star)
REL_TOL = 1e-4       # eigenvalue-ratio columns, relative, against the float64 reference
ABS_FLOOR = 1e-9     # below this the reference's own ratios are rounding noise (exactly collinear /
                     # coplanar voxel sets give +-1e-13 instead of 0)


def assert_features_close(got, ref, radii, ncol=4):
    """population exact; centroid abs <= 1e-4 * r; ratio columns rel <= 1e-4 (+ noise floor)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    for s, r in enumerate(radii):
        base = s * ncol
        assert np.array_equal(got[:, base], ref[:, base]), "population differs at scale %d" % s
        cen = np.abs(got[:, base + 1] - ref[:, base + 1])
        assert cen.max(initial=0.0) <= 1e-4 * r, "centroid off by %g at scale %d" % (cen.max(), s)
        for j in (2, 3):
            d = np.abs(got[:, base + j] - ref[:, base + j])
            lim = REL_TOL * np.abs(ref[:, base + j]) + ABS_FLOOR
            bad = d > lim
            assert not bad.any(), "ratio col %d scale %d: max err %g (ref %g)" % (
                j, s, d[bad].max(), ref[:, base + j][bad][np.argmax(d[bad])])
