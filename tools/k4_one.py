"""One short run of the exact 5-NN kernel (for ncu): 8 M-point plane map, 1 M queries in voxel order (as the frame loop presents them) or random order (K4_ORDER=random)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from pf_loader import pfb
import sweep
M, nq = int(os.environ.get("K4_POINTS", "8000000")), 1_000_000
xyz = sweep.plane_map(M)[:M]
rng = np.random.default_rng(4001)
q = np.zeros((nq, 4), np.float32)
q[:, :3] = xyz[rng.integers(0, M, nq)] + rng.normal(0, 0.2, (nq, 3)).astype(np.float32)
if os.environ.get("K4_ORDER", "voxel") == "voxel":
    vk = np.floor(q[:, :3] / np.float32(0.4)).astype(np.int64)
    q = np.ascontiguousarray(q[np.lexsort((vk[:, 0], vk[:, 1], vk[:, 2]))])
idx, d2, mb, mq = pfb.capi.knn5_timed(pfb.capi.make_points(xyz, r=0, g=1), q, reps=3)
print("map", M, "queries", nq, "build ms", mb, "query ms", mq, "Gq/s", nq / mq / 1e6, "valid", float((idx[:, 4] >= 0).mean()))
