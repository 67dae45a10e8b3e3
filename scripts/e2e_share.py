"""dev tool: float64 host-buffer call with the three routes of a pinned result (NBR_HOST_WIDEN = auto / host / device)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for mode in ("auto", "host", "device", "auto"):
    e = dict(os.environ); e["NBR_HOST_WIDEN"] = mode; e["NBR_HOST_STATS"] = "1"
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts/e2e_sweep.py"), "child", "widen " + mode], env=e)
