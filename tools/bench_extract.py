"""Micro-benchmark of the batched extraction kernels with HBM-resident inputs (K1, SURVEY.md section 8 D4: 32 B/point)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pf_loader import pfb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--distinct", type=int, default=8)
    a = ap.parse_args()
    capi = pfb.capi
    p = pfb.synth.config("cfg2")
    scans = [pfb.synth.scan(p, f) for f in range(a.distinct)]
    stride = 115200
    ex = capi.Extractor(num_lines=64, max_points=stride, max_batch=a.batch)
    x = np.zeros((a.batch, stride, 4), np.float32)
    n = np.zeros(a.batch, np.int32)
    for i in range(a.batch):
        s = scans[i % a.distinct]
        x[i, :len(s)] = s
        n[i] = len(s)
    dev = torch.device("cuda:0")
    dx = torch.from_numpy(x).to(dev)
    dn = torch.from_numpy(n).to(dev)
    dedge = torch.empty((a.batch, ex.edge_stride, 4), dtype=torch.float32, device=dev)
    dsurf = torch.empty((a.batch, stride, 4), dtype=torch.float32, device=dev)
    dne = torch.zeros(a.batch, dtype=torch.int32, device=dev)
    dns = torch.zeros(a.batch, dtype=torch.int32, device=dev)
    stream = torch.cuda.ExternalStream(ex.stream)
    torch.cuda.synchronize()

    def go():
        ex.run_batch_device(dx.data_ptr(), dn.data_ptr(), a.batch, stride, dedge.data_ptr(), dne.data_ptr(), dsurf.data_ptr(),
                            dns.data_ptr(), 0)
    for _ in range(3):
        go()
    ex.sync()
    times = []
    for _ in range(a.iters):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            go()
            e1.record(stream)
        ex.sync()
        times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    pts = int(n.sum())
    out_pts = int(dne.sum().item() + dns.sum().item())
    alg = 16 * pts + 16 * out_pts
    print(json.dumps({"batch": a.batch, "ms": ms, "scans_per_s": a.batch / ms * 1e3, "points": pts, "out_points": out_pts,
                      "algorithmic_GBps": alg / ms / 1e6, "sb_32B_per_pt_GBps": 32 * pts / ms / 1e6, "times": times}))


if __name__ == "__main__":
    main()
