"""Steady-state per-frame kernel breakdown from an ncu launch list of tools/profile_frames.py (last N frames)."""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lines = [l for l in open(path) if l.startswith('"')]
rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
ids = [i for i, r in enumerate(rows) if "k_predict" in r["Kernel Name"]]
lo, hi = ids[-nf - 1] - 4, ids[-1] - 4     # a frame starts 4 launches before k_predict (set_int, classify, index, extract)
sel = rows[lo:hi]
agg = OrderedDict()
for r in sel:
    k = r["Kernel Name"].split("(")[0].replace("pf::", "").replace("<unnamed>::", "")
    v = float(r["Metric Value"].replace(",", "")) / 1000.0
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{nf} frames, {len(sel)} launches, {tot / nf:.1f} us/frame kernel time (ncu: serialised, cold caches), {len(sel) / nf:.1f} launches/frame")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / nf:8.1f} us/frame {100 * us / tot:5.1f}%  n/frame={n / nf:4.1f} avg {us / n:6.2f}  {k}")
