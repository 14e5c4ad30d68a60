"""Summarise the per-source-line stall samples of one kernel from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
fname = ""
agg = {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 6 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        v = float(r[hdr.index("# Samples")])
        ins = float(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    key = (fname, r[0], r[1].strip()[:120])
    a = agg.setdefault(key, [0.0, 0.0])
    a[0] += v
    a[1] += ins
tot = sum(v[0] for v in agg.values()) or 1.0
toti = sum(v[1] for v in agg.values()) or 1.0
data = sorted(agg.items(), key=lambda kv: -kv[1][0])
print("samples%  inst%   file:line  source")
for (f, l, s), (v, ins) in data[:top]:
    print("%5.1f%%  %5.1f%%  %s:%s  %s" % (100 * v / tot, 100 * ins / toti, f, l, s))
