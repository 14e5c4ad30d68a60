"""K9 / K4 legs of bench.py alone (for quick iteration and ncu captures)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pf_loader import pfb  # noqa: E402

print(json.dumps(bench.extra_kernel_legs(pfb.capi, 0)))
