"""
golden vectors for the tile + halo selection rule, produced by running the UNMODIFIED reference
(nimrud/utils/geometry.py:203-253, nested_regions) in the build container:

    PYTHONPATH=/root/reference python tests/golden/make_golden_regions.py

the inputs follow the reference's own test (utils/tests/geometry_tests.py:353-389): a unit-cube query set, a
search space in [-1, 2]^3, the region [0.25, 0.75]^3 with buffer radius 0.5, and a region that holds no point.
only the seeds and the index arrays are stored; the test regenerates the clouds.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def clouds():
    rs = np.random.RandomState(10)
    query = rs.rand(5000, 3)
    search = rs.rand(20000, 3) * 3 - 1
    # a few points exactly on the region's faces: the rule is inclusive on both sides
    search[:3] = [[-0.25, 0.5, 0.5], [0.5, 1.25, 0.5], [0.5, 0.5, -0.25]]
    query[:2] = [[0.25, 0.25, 0.25], [0.75, 0.75, 0.75]]
    return query.astype(np.float32).astype(np.float64), search.astype(np.float32).astype(np.float64)


def main():
    sys.path.insert(0, "/root/reference")
    from nimrud.utils import geometry          # the reference itself; only needed to (re)generate the fixture
    query, search = clouds()
    lo, hi = np.array([0.25, 0.25, 0.25]), np.array([0.75, 0.75, 0.75])
    qi, si = geometry.nested_regions(query, search, 0.5, lo, hi)
    lo2 = np.ones(3) * 100
    qe, se = geometry.nested_regions(query, search, 0.5, lo2, lo2 + 10)
    assert qe.size == 0 and se.size == 0
    np.savez_compressed(os.path.join(HERE, "golden_regions.npz"), query_idx=qi.astype(np.int32),
                        search_idx=si.astype(np.int32), lo=lo, hi=hi, buffer=np.float64(0.5),
                        empty_query_idx=qe.astype(np.int32), empty_search_idx=se.astype(np.int32))
    print("query", qi.size, "search", si.size)


if __name__ == "__main__":
    main()
