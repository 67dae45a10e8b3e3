"""dev tool: float64 host-buffer call with different shares of device-widened rows (NBR_HOST_DIRECT_SHARE)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for share in ("0", "0.05", "0.1", "0.15", "0.2", "0.1", "0"):
    e = dict(os.environ); e["NBR_HOST_DIRECT_SHARE"] = share
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts/e2e_sweep.py"), "child", "share " + share], env=e)
