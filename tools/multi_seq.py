"""Throughput of S independent sequences run concurrently on one GPU (one handle pair + stream set each, one host thread serving
them in turn).  usage: multi_seq.py [S] [frames]   (try CUDA_DEVICE_MAX_CONNECTIONS=32: more hardware queues than the default 8)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from concurrent.futures import ThreadPoolExecutor
from pf_loader import pfb
capi = pfb.capi
S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
K = int(sys.argv[2]) if len(sys.argv) > 2 else 100
seqs = []
for i in range(S):
    p = pfb.synth.config(f"cfg5.{i % 8}")
    pfb.synth.scan(p, 0)
    with ThreadPoolExecutor(16) as tp:
        scans = list(tp.map(lambda f: pfb.synth.scan(p, f), range(K)))
    pin = []
    for sc in scans:
        a, ptr = capi.pinned_array((len(sc), 4), np.float32)
        a[:] = sc
        pin.append(a)
    seqs.append(pin)
for rep in range(3):
    H = [(capi.Extractor(num_lines=64, max_points=115200), capi.Odometry(0.4, 0, 0.4, 75, max_map_points=1 << 19, max_features=115200)) for _ in range(S)]
    t0 = time.perf_counter()
    t_sub = t_wait = 0.0
    prev = [capi.frame_submit(H[i][0], H[i][1], seqs[i][0]) for i in range(S)]
    for k in range(1, K):
        a = time.perf_counter()
        cur = [capi.frame_submit(H[i][0], H[i][1], seqs[i][k]) for i in range(S)]
        b = time.perf_counter()
        for i in range(S):
            capi.frame_wait(H[i][1], prev[i])
        c = time.perf_counter()
        t_sub += b - a; t_wait += c - b
        prev = cur
    for i in range(S):
        capi.frame_wait(H[i][1], prev[i])
    dt = time.perf_counter() - t0
    print(f"   host time inside frame_submit {100 * t_sub / dt:.0f} %, inside frame_wait {100 * t_wait / dt:.0f} % of the wall time "
          f"({1e6 * t_sub / (S * (K - 1)):.1f} / {1e6 * t_wait / (S * (K - 1)):.1f} us per frame)")
    for ex, od in H:
        ex.close(); od.close()
    print(f"S={S} frames={K} rep {rep}: {S * K / dt:.0f} scans/s  (CUDA_DEVICE_MAX_CONNECTIONS={os.environ.get('CUDA_DEVICE_MAX_CONNECTIONS', 'default')})", flush=True)
