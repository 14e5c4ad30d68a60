"""Stress check of the queued frame pipeline (down-sampling overlap, graph re-capture as the maps grow): N frames of a configuration
through pf_frame_submit / pf_frame_wait with several frames in flight must give the poses of blocking pf_frame_process calls, bit for bit.
usage: stress_pipeline.py [cfg] [frames] [in_flight]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf_loader import pfb  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 300
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 4
capi = pfb.capi
p = pfb.synth.config(cfg)
prm = {"cfg4": (0.4, 0, 1.0, 200)}.get(cfg, (0.4, 0, 0.4, 75))
scans = [pfb.synth.scan(p, f) for f in range(frames)]


def make():
    return capi.Extractor(num_lines=p.sensor_lines, max_points=131072), capi.Odometry(*prm, max_map_points=1 << 21)


ex, od = make()
t0 = time.perf_counter()
sync = np.array([capi.frame_process(ex, od, s) for s in scans])
t_sync = time.perf_counter() - t0
st = od.stats()
ex.close(); od.close()
ex, od = make()
t0 = time.perf_counter()
ids, out = [], []
for k, s in enumerate(scans):
    ids.append(capi.frame_submit(ex, od, s))
    if k >= depth:
        out.append(capi.frame_wait(od, ids[k - depth]))
out += [capi.frame_wait(od, i) for i in ids[-depth:]]
t_pipe = time.perf_counter() - t0
same = np.array(out).tobytes() == sync.tobytes()
print(f"{cfg}: {frames} frames, maps {st['map_edge']} / {st['map_surf']}, blocking {frames / t_sync:.0f} scans/s, queued {frames / t_pipe:.0f} scans/s, "
      f"graph captures {od.graph_captures if hasattr(od, 'graph_captures') else '?'}, identical: {same}")
sys.exit(0 if same else 1)
