"""Global map (K10) at scale: grow the map with random points, then time steady-state updates (wall clock around
pf_mapping_update + pf_mapping_map_size, which waits for the update)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf_loader import pfb  # noqa: E402

capi = pfb.capi
target = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
rng = np.random.default_rng(7)
m = capi.Mapping(0.4, max_map_points=target + 2_000_000, max_points=262144)
rt = np.eye(4)[:3].reshape(12)
while m.size() < target:
    pts = ((rng.random((262144, 4), dtype=np.float32) - 0.5) * np.array([240, 240, 60, 1], np.float32)).astype(np.float32)
    m.update(pts, rt)
n0 = m.size()
pts = ((rng.random((80000, 4), dtype=np.float32) - 0.5) * np.array([120, 120, 10, 1], np.float32)).astype(np.float32)
ts = []
for k in range(8):
    t0 = time.perf_counter()
    m.update(pts + np.float32(0.01 * k), rt)
    n1 = m.size()
    ts.append(time.perf_counter() - t0)
ms = 1e3 * float(np.median(ts[2:]))
print({"map_points": n0, "map_points_after": n1, "new_points": len(pts), "ms_per_update": ms,
       "algorithmic_gbs": 40.0 * n0 / (ms * 1e-3) / 1e9, "stats": m.stats()})
