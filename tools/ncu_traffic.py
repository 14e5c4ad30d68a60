"""DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) per launch of the roofline kernels, from ncu --set full captures;
writes profiles/traffic_r2.json, which bench.py reports as roofline.traffic.
usage: ncu_traffic.py k1.ncu-rep k9.ncu-rep k4.ncu-rep"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernels(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        def val(name):
            v = float(d[name].replace(",", ""))
            u = units[hdr.index(name)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        res.append({"kernel": d["Kernel Name"].split("(")[0].split("<")[0].split("::")[-1].replace("void ", "").strip(), "dram_read": val("dram__bytes_read.sum"),
                    "dram_write": val("dram__bytes_write.sum"), "time_us": float(d["gpu__time_duration.sum"].replace(",", "")) / (1e3 if units[hdr.index("gpu__time_duration.sum")] in ("ns", "nsecond") else 1),
                    "l2_hit_pct": float(d["lts__t_sector_hit_rate.pct"])})
    return res


def first(ks, name):
    return next(k for k in ks if k["kernel"] == name)


k1, k9, k4 = kernels(sys.argv[1]), kernels(sys.argv[2]), kernels(sys.argv[3])
out = {"source": [os.path.basename(a) for a in sys.argv[1:4]],
       "note": "ncu --set full --clock-control none; caches flushed before every kernel, so re-reads that hit L2 in the real pipeline show up as DRAM here"}
c, x, e = first(k1, "k_ring_classify"), first(k1, "k_ring_index"), first(k1, "k_sector_extract")
t_all = c["time_us"] + x["time_us"] + e["time_us"]
out["k1"] = {"classify": c, "index": x, "extract": e,
             "bytes_per_launch_group": sum(k["dram_read"] + k["dram_write"] for k in (c, x, e)),
             "share_of_group_time": {"k_ring_classify": c["time_us"] / t_all, "k_ring_index": x["time_us"] / t_all,
                                     "k_sector_extract": e["time_us"] / t_all}}
a, b = first(k9, "k_mm_count"), first(k9, "k_mm_write")
out["k9"] = {"count": a, "write": b, "bytes_per_update": a["dram_read"] + a["dram_write"] + b["dram_read"] + b["dram_write"]}
q = first(k4, "k_knn5_tap")
out["k4"] = {"knn": q, "bytes_per_launch": q["dram_read"] + q["dram_write"]}
with open(os.path.join(ROOT, "profiles", "traffic_r2.json"), "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(out, indent=1))
