"""Runs a few frames of the synthetic sequence through the frame pipeline (for ncu launch lists / captures)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf_loader import pfb  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--cfg", default="cfg2")
a = ap.parse_args()
capi = pfb.capi
p = pfb.synth.config(a.cfg)
scans = [pfb.synth.scan(p, f) for f in range(a.frames)]
ex = capi.Extractor(num_lines=p.sensor_lines, max_points=115200)
od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19, max_features=115200)
t0 = time.perf_counter()
for s in scans:
    pose = capi.frame_process(ex, od, s)
dt = time.perf_counter() - t0
print("frames", a.frames, "ms/frame", 1e3 * dt / a.frames, "pose", np.round(pose, 4), "launches", ex.launches + od.launches, od.stats())
if os.environ.get("PF_ODOM_TIMING"):
    import ctypes as C
    ms = (C.c_float * 5)()
    acc = np.zeros(5)
    ex2 = capi.Extractor(num_lines=p.sensor_lines, max_points=115200)
    od2 = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19, max_features=115200)
    tt = []
    for k, s in enumerate(scans):
        t0 = time.perf_counter()
        capi.frame_process(ex2, od2, s)
        tt.append(time.perf_counter() - t0)
        if k >= 12:
            capi.check(capi.lib().pf_odom_get_phase_ms(od2.h, ms))
            acc += np.array(ms[:])
    print("steady-state phase ms (avg): downsample %.3f grid %.3f passes-to-last-assoc %.3f last-5-evals %.3f map-update %.3f" % tuple(acc / (len(scans) - 12)),
          " wall ms/frame (last 4): %.3f" % (1e3 * np.mean(tt[12:])))
