"""One short run of the grid-wide normal-equation kernel (for ncu): 20 M residual blocks, 3 launches."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from pf_loader import pfb
import sweep
n = int(os.environ.get("K7_BLOCKS", "10000000"))
e, s = sweep.residual_blocks(n)
H, g, c, ms = pfb.capi.eval_normal_eq_timed(sweep.SWEEP_POSE, e, s, reps=3)
print("blocks", 2 * n, "ms", ms, "GB/s", 128.0 * n / ms / 1e6, "cost", c)
