"""Map-size sweep of BASELINE.json configs[4]: M in {1, 2, 5, 10, 20, 50} x 10^6 map points / residual blocks for
  K3  search-grid build (stands in for KdTreeFLANN::setInputCloud, /root/reference/src/odomEstimationClass.cpp:249-250),
  K4  exact 5-NN queries (nearestKSearch, :299, :447), 10^6 queries near map points (sigma 0.2 m), in random order and in the
      voxel order the frame loop presents them in,
  K7  residual + Jacobian + Huber + J^T J over M residual blocks (src/lidarOptimization.cpp:12-78), the grid-wide streaming kernel,
with the CPU kd-tree (oracle restatement of FLANN's KDTreeSingleIndex, leaf 15) timed beside K3 / K4 at the sizes it finishes in
seconds.  Data per SURVEY.md section 8 D2: voxel-centroid-like points on random planes in a +-100 m cube, seed 4000.

usage: sweep.py [--sizes 1,2,5,10,20,50] [--out profiles/x2_sweep.json]          (also imported by bench.py)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def plane_map(m_points, seed=4000, leaf=0.4, half=100.0):
    """~one point per occupied `leaf` voxel on random planes through the cube; returns (n, 3) float32, n >= m_points."""
    rng = np.random.default_rng(seed)
    out, total = [], 0
    k = int(2 * half / leaf)
    ii, jj = np.meshgrid(np.arange(-k, k, dtype=np.float32), np.arange(-k, k, dtype=np.float32), indexing="ij")
    ii, jj = ii.ravel() * np.float32(leaf), jj.ravel() * np.float32(leaf)
    while total < m_points:
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        u = np.cross(n, [0.3, 0.5, 0.8]); u /= np.linalg.norm(u)
        v = np.cross(n, u)
        o = rng.uniform(-0.6 * half, 0.6 * half, 3)
        p = (o[None, :] + ii[:, None] * u[None, :].astype(np.float32) + jj[:, None] * v[None, :].astype(np.float32)).astype(np.float32)
        p = p[(np.abs(p) < half - 0.5).all(1)]
        p += rng.uniform(-0.12, 0.12, p.shape).astype(np.float32)      # centroid-like: off the lattice
        out.append(p)
        total += len(p)
    return np.concatenate(out)


def residual_blocks(n, seed=4100):
    """n edge blocks [p, a, b] and n surf blocks [p, n, d] (float64), a mix of inlier and Huber-range residuals."""
    rng = np.random.default_rng(seed)
    base = min(n, 1 << 20)
    p = rng.uniform(-50, 50, (base, 3))
    u = rng.normal(size=(base, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    off = rng.normal(0, 0.06, (base, 3))
    a = p + off + 0.1 * u
    b = p + off - 0.1 * u
    edge = np.concatenate([p, a, b], 1)
    nn = rng.normal(size=(base, 3)); nn /= np.linalg.norm(nn, axis=1, keepdims=True)
    d = -(nn * p).sum(1) + rng.normal(0, 0.06, base)
    surf = np.concatenate([p, nn, d[:, None]], 1)
    reps = (n + base - 1) // base
    return np.ascontiguousarray(np.tile(edge, (reps, 1))[:n]), np.ascontiguousarray(np.tile(surf, (reps, 1))[:n])


SWEEP_POSE = np.array([0.01, -0.02, 0.015, 0.0, 0.05, -0.03, 0.02])
SWEEP_POSE[3] = np.sqrt(1 - (SWEEP_POSE[:3] ** 2).sum())


def run_sweep(capi, sizes_m=(1, 2, 5, 10, 20, 50), nq=1_000_000, device=0, oracle=None, cpu_sizes_m=(1, 5), cpu_queries=100_000, peak=None):
    if peak is None:
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            peak = 6650.0
    t0 = time.perf_counter()
    xyz_all = plane_map(int(max(sizes_m) * 1e6))
    edge_all, surf_all = residual_blocks(int(max(sizes_m) * 1e6) // 2)
    t_gen = time.perf_counter() - t0
    pose = SWEEP_POSE
    rows = []
    rng = np.random.default_rng(4001)
    for sm in sizes_m:
        M = int(sm * 1e6)
        pts = capi.make_points(xyz_all[:M], r=0, g=1)
        sel = rng.integers(0, M, nq)
        q = np.zeros((nq, 4), np.float32)
        q[:, :3] = xyz_all[sel] + rng.normal(0, 0.2, (nq, 3)).astype(np.float32)
        idx, d2, ms_build, ms_q_rand = capi.knn5_timed(pts, q, reps=3, device=device)
        # the frame loop hands the queries over in voxel order (VoxelGrid output: z, y, x ascending at leaf 0.4)
        vk = np.floor(q[:, :3] / np.float32(0.4)).astype(np.int64)
        order = np.lexsort((vk[:, 0], vk[:, 1], vk[:, 2]))
        qs = np.ascontiguousarray(q[order])
        idx_s, d2_s, _, ms_q_sorted = capi.knn5_timed(pts, qs, reps=3, device=device)
        assert np.array_equal(idx_s, idx[order]), "query order changed the k-NN result"
        ne = M // 2
        H, g, cost, ms_k7 = capi.eval_normal_eq_timed(pose, edge_all[:ne], surf_all[:ne], reps=4, device=device)
        bytes_k7 = 72.0 * ne + 56.0 * ne
        row = {"map_points": M, "queries": nq,
               "k3_grid_build_ms": ms_build, "k3_gbs": 36.0 * M / (ms_build * 1e-3) / 1e9, "k3_frac": 36.0 * M / (ms_build * 1e-3) / 1e9 / peak,
               "k4_queries_per_s_random_order": nq / (ms_q_rand * 1e-3), "k4_queries_per_s_voxel_order": nq / (ms_q_sorted * 1e-3),
               "k4_algorithmic_gbs_voxel_order": 136.0 * nq / (ms_q_sorted * 1e-3) / 1e9,
               "k4_valid_fraction": float((idx[:, 4] >= 0).mean()),
               "k7_residual_blocks": 2 * ne, "k7_ms": ms_k7, "k7_gbs": bytes_k7 / (ms_k7 * 1e-3) / 1e9, "k7_frac": bytes_k7 / (ms_k7 * 1e-3) / 1e9 / peak,
               "k7_blocks_per_s": 2 * ne / (ms_k7 * 1e-3), "k7_cost": cost}
        if oracle is not None and sm in cpu_sizes_m:
            ci, cd, s_build, s_query = oracle.knn5_timed(pts, q[:cpu_queries])
            row["cpu_kdtree"] = {"build_s": s_build, "queries": cpu_queries, "queries_per_s": cpu_queries / s_query, "cores": 1,
                                 "identical_to_gpu": bool(np.array_equal(ci, idx[:cpu_queries]) and np.array_equal(cd, d2[:cpu_queries]))}
            # CPU J^T J on a bounded sample of the same blocks
            nb = min(ne, 500_000)
            t1 = time.perf_counter()
            oracle.eval_normal_eq(pose, edge_all[:nb], surf_all[:nb])
            dt = time.perf_counter() - t1
            row["cpu_normal_eq"] = {"blocks": 2 * nb, "blocks_per_s": 2 * nb / dt, "cores": 1}
        rows.append(row)
        del pts
    return {"data": "random planes in a +-100 m cube, ~one point per 0.4 m voxel (seed 4000); queries = map points + N(0, 0.2 m); residual blocks: "
                    "half point-to-line (72 B), half point-to-plane (56 B), base set of 2^20 tiled", "generation_s": t_gen, "peak_gbs": peak,
            "bytes": {"k3": "36 B per map point (16 read + 20 written)", "k4": "136 B per query", "k7": "72 B edge / 56 B surf per residual block"},
            "rows": rows}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,2,5,10,20,50")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from pf_loader import pfb
    import oracle as O
    res = run_sweep(pfb.capi, tuple(float(x) for x in a.sizes.split(",")), oracle=O)
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        open(os.path.join(ROOT, a.out), "w").write(txt + "\n")
