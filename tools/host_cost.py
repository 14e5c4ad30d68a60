"""Host-side cost of the frame calls: wall time spent inside pf_frame_submit / pf_frame_wait per frame (single sequence, frames queued
one ahead).  If submit costs ~150 us per frame, S sequences served by one driver context cannot exceed ~6.7 k scans/s however idle the GPU."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pf_loader import pfb
capi = pfb.capi
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
p = pfb.synth.config("cfg5.0")
pin = []
for f in range(K):
    sc = pfb.synth.scan(p, f)
    a, _ = capi.pinned_array((len(sc), 4), np.float32)
    a[:] = sc
    pin.append(a)
for rep in range(3):
    ex = capi.Extractor(num_lines=64, max_points=115200)
    od = capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19, max_features=115200)
    ts, tw = [], []
    t00 = time.perf_counter()
    prev = capi.frame_submit(ex, od, pin[0])
    for k in range(1, K):
        t0 = time.perf_counter()
        cur = capi.frame_submit(ex, od, pin[k])
        t1 = time.perf_counter()
        capi.frame_wait(od, prev)
        t2 = time.perf_counter()
        ts.append(t1 - t0); tw.append(t2 - t1)
        prev = cur
    capi.frame_wait(od, prev)
    dt = time.perf_counter() - t00
    ts, tw = np.array(ts[20:]), np.array(tw[20:])
    print(f"rep {rep}: {K / dt:.0f} scans/s; submit {1e6 * ts.mean():.1f} us (median {1e6 * np.median(ts):.1f}), wait {1e6 * tw.mean():.1f} us per frame (steady state)", flush=True)
    ex.close(); od.close()
