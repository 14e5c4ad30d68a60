"""K1 only: batched extraction roofline leg of bench.py (for quick iteration and ncu captures)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from pf_loader import pfb  # noqa: E402

p = pfb.synth.config("cfg2")
scans = [pfb.synth.scan(p, f) for f in range(8)]
r = bench.roofline_leg(pfb, pfb.capi, torch, scans, torch.device("cuda:0"))
print(json.dumps(r))
