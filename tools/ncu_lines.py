"""Per-source-line instruction counts / stall samples of one kernel from an .ncu-rep captured with --import-source on.
usage: ncu_lines.py report.ncu-rep kernel_substring [top_n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg, tot, tot_s = None, None, {}, 0, 0
seen_fn = set()
active = False
for r in rows:
    if not r:
        continue
    if r[0] in ("Function Name", "Kernel Name"):
        active = kern in r[1] and ("k", r[1]) not in seen_fn
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if not active or hdr is None or len(r) < len(hdr) - 2:
        continue
    d = dict(zip(hdr, r))
    if r[0]:   # a source line row (aggregated over its SASS)
        try:
            n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
        except ValueError:
            continue
        key = (int(r[0]), r[1].strip()[:100])
        a = agg.setdefault(key, [0, 0])
        a[0] += n; a[1] += s
        tot += n; tot_s += s
print(f"kernel ~{kern}: {tot} warp instructions, {tot_s} samples")
for (ln, src), (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*n/max(tot,1):5.1f}% inst {100*s/max(tot_s,1):5.1f}% smp  L{ln:4d}  {src}")
if len(sys.argv) > 4:
    # extra args: line ranges "name:lo-hi"
    print("-- by range")
    for spec in sys.argv[4:]:
        name, rng = spec.split(":")
        lo, hi = map(int, rng.split("-"))
        n = sum(v[0] for (ln, _), v in agg.items() if lo <= ln <= hi)
        s = sum(v[1] for (ln, _), v in agg.items() if lo <= ln <= hi)
        print(f"{name:14s} L{lo}-{hi}: {100*n/max(tot,1):5.1f}% inst  {100*s/max(tot_s,1):5.1f}% samples")
