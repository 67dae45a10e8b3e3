"""
Golden vectors for the classifier hand-off metrics, produced by RUNNING THE REFERENCE's own
functions (nimrud/prototypes/ml.py: mc_confusion :521, user_producer :465, three_metrics :491).

    python tests/golden/make_golden_learning.py

ml.py imports matplotlib at module level (absent in this image) although none of the three
functions uses it, so an empty stand-in module is registered before the import; nothing else of
the reference is touched.  writes golden_learning.npz.
"""
import os
import sys
import types

import numpy as np

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, "/root/reference")
from nimrud.prototypes import ml as ref_ml      # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    rs = np.random.RandomState(7)
    out = {}
    for case, (n, classes) in enumerate(((5000, 4), (1200, 6), (300, 2))):
        truth = np.repeat(np.arange(classes), n // classes)              # balanced, as three_metrics assumes
        lies = truth.copy()
        flip = rs.rand(len(truth)) < 0.3
        lies[flip] = rs.randint(0, classes, flip.sum())
        conf = ref_ml.mc_confusion(lies, truth)
        user, prod = ref_ml.user_producer(conf)
        out["truth%d" % case] = truth
        out["lies%d" % case] = lies
        out["conf%d" % case] = conf
        out["user%d" % case] = user
        out["prod%d" % case] = prod
        out["three%d" % case] = ref_ml.three_metrics(conf)
    np.savez_compressed(os.path.join(HERE, "golden_learning.npz"), **out)
    print("wrote golden_learning.npz")


if __name__ == "__main__":
    main()
