"""Repro helper: cfg4 frames through the queued pipeline with given capacities.  dbg_cfg4.py max_points max_features map_cap frames [cfg]"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf_loader import pfb
capi = pfb.capi
mp, mf, mc, nf = (int(x) for x in sys.argv[1:5])
cfg = sys.argv[5] if len(sys.argv) > 5 else "cfg4"
p = pfb.synth.config(cfg)
from concurrent.futures import ThreadPoolExecutor
pfb.synth.scan(p, 0)
with ThreadPoolExecutor(16) as tp:
    scans = list(tp.map(lambda f: pfb.synth.scan(p, f), range(nf)))
ex = capi.Extractor(num_lines=p.sensor_lines, max_points=mp)
od = capi.Odometry(0.4, 0, 1.0, 200, max_map_points=mc, max_features=mf)
prev = capi.frame_submit(ex, od, scans[0])
k = 0
try:
    for k in range(1, nf):
        fid = capi.frame_submit(ex, od, scans[k])
        capi.frame_wait(od, prev)
        prev = fid
        if k % 100 == 0:
            print(k, od.stats(), flush=True)
except Exception as e:
    print("FAILED at frame", k, e, flush=True)
    sys.exit(1)
print("ok", capi.frame_wait(od, prev), od.stats())
