#!/bin/bash
# dev tool: build the fused kernel with its shared-memory bounds checks on and run the small all-kernels script plus a
# 2M-point scene through it; prints the number of violations (expected: 0)
set -e
bash scripts/variant.sh bounds rows3.cu -DNBR_BOUNDS_CHECK=1
NIMRUD_B200_LIB=nimrud_b200/lib/variants/bounds.so python - <<'PY'
import runpy, sys, numpy as np, torch
sys.argv = ["sanitize_small.py"]
runpy.run_path("scripts/sanitize_small.py")
from nimrud_b200 import _lib, multiscale, synth
cloud = synth.urban_scene(2_000_000, seed=3, device="cuda")
multiscale.process_single_core(cloud, cloud, [0.1, 0.2, 0.4, 0.8, 1.6], [0.3, 0.6, 1.2, 2.4, 4.8], out_dtype=np.float32)
multiscale.process_single_core(cloud[:500_000].contiguous(), cloud, [0.2, 0.2], [0.4, 0.7], out_dtype=np.float64)
torch.cuda.synchronize()
print("shared-memory bounds violations in rows3_kernel:", _lib.lib().nbr_debug_bounds_violations())
PY
