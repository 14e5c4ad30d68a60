"""Pinning of oracle B (the restated odometry) as far as this image allows -- test infrastructure, used by tests/ and run as a tool
to write profiles/pinning_r2.json.  The reference's odometry cannot be built here (PCL / FLANN / Eigen / Ceres are absent), so the
third-party numerics the oracle restates are tied to INDEPENDENT implementations instead:

  lm_fixed_point   the oracle's Levenberg-Marquardt state machine (Ceres restated, /root/reference/src/odomEstimationClass.cpp:254-271),
                   run to convergence, against scipy.optimize.least_squares(loss='huber', f_scale=0.1) on the same residual blocks
                   with residual functions written in numpy (nothing shared with the oracle).
  flip_report      the two threshold decisions of the association pass -- line fit lambda_2 > 3 lambda_1 (:326) and plane fit
                   |n.p + d| <= 0.2 for all five neighbours (:469-471) -- from the oracle's Jacobi eigen-solver / column-pivoted
                   Householder QR against numpy.linalg.eigh / pinv (LAPACK), and against the GPU kernels.
  noise_floor      how far the conventions the reference leaves open move a trajectory: surf emission order (ascending curvature in the
                   reference, ring position in ours) and the point order inside a voxel (unstable std::sort in the reference, stable
                   in ours).  This is the floor under which trajectory differences carry no information.

usage: pinning.py [--frames 100] [--gpu] [--out profiles/pinning_r2.json]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

PT = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("a", "u1")])


def to_points(xyz4):
    out = np.zeros(len(xyz4), PT)
    out["x"], out["y"], out["z"] = xyz4[:, 0], xyz4[:, 1], xyz4[:, 2]
    out["a"] = 255
    return out


def quat_to_R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def rodrigues(w):
    th = np.linalg.norm(w)
    if th < 1e-12:
        return np.eye(3)
    k = w / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def transform_queries(pose, q_pts):
    """pointAssociateToMap (:162-168): double transform, float result."""
    R = quat_to_R(pose[:4])
    p = np.stack([q_pts["x"], q_pts["y"], q_pts["z"]], 1).astype(np.float64)
    w = p @ R.T + pose[4:]
    out = np.zeros((len(p), 4), np.float32)
    out[:, :3] = w.astype(np.float32)
    return out


# ---------------------------------------------------------------------------------------------------------------------------
# residual sets of consecutive frame pairs
# ---------------------------------------------------------------------------------------------------------------------------
def frame_pair(pfb, O, k, cfg="cfg2"):
    """map = raw features of frame k (what initMapWithPoints would hold), queries = VoxelGrid(features of frame k + 1)."""
    p = pfb.synth.config(cfg)
    s0, s1 = pfb.synth.scan(p, k), pfb.synth.scan(p, k + 1)
    r0, r1 = O.extract(s0, num_lines=p.sensor_lines, order=1), O.extract(s1, num_lines=p.sensor_lines, order=1)
    maps = [to_points(s0[r0["edge_idx"]]), to_points(s0[r0["surf_idx"]])]
    qs = [O.voxel_downsample(to_points(s1[r1["edge_idx"]]), 0.4), O.voxel_downsample(to_points(s1[r1["surf_idx"]]), 0.8)]
    g0, g1 = pfb.synth.pose(p, k), pfb.synth.pose(p, k + 1)
    R0 = quat_to_R(g0[:4])
    R01 = R0.T @ quat_to_R(g1[:4])
    t01 = R0.T @ (g1[4:] - g0[4:])
    # quaternion of R01 (rotation about z only in the generator)
    yaw = np.arctan2(R01[1, 0], R01[0, 0])
    pose = np.array([0, 0, np.sin(yaw / 2), np.cos(yaw / 2), *t01])
    return maps, qs, pose


def residual_set(O, maps, qs, pose, params=(0, 0.4, 75)):
    e9, s7 = np.zeros((0, 9)), np.zeros((0, 7))
    for kind in (0, 1):
        _, _, flag, geom = O.associate(kind, maps[kind], qs[kind], pose, *params)
        sel = flag == 2
        p = np.stack([qs[kind]["x"], qs[kind]["y"], qs[kind]["z"]], 1).astype(np.float64)[sel]
        if kind == 0:
            e9 = np.concatenate([p, geom[sel, :6]], 1)
        else:
            s7 = np.concatenate([p, geom[sel, :4]], 1)
    return e9, s7


# ---------------------------------------------------------------------------------------------------------------------------
# (i) LM fixed point against scipy
# ---------------------------------------------------------------------------------------------------------------------------
def numpy_residuals(R, t, e9, s7):
    """src/lidarOptimization.cpp:18-24 (edge) and :60-61 (surf), written independently of the oracle."""
    out = []
    if len(e9):
        lp = e9[:, :3] @ R.T + t
        nu = np.cross(lp - e9[:, 3:6], lp - e9[:, 6:9])
        out.append(np.linalg.norm(nu, axis=1) / np.linalg.norm(e9[:, 3:6] - e9[:, 6:9], axis=1))
    if len(s7):
        lp = s7[:, :3] @ R.T + t
        out.append((s7[:, 3:6] * lp).sum(1) + s7[:, 6])
    return np.concatenate(out) if out else np.zeros(0)


def scipy_fixed_point(e9, s7, pose0):
    from scipy.optimize import least_squares
    R0, t0 = quat_to_R(pose0[:4]), np.asarray(pose0[4:], float)

    def f(d):
        Rd = rodrigues(d[:3])
        return numpy_residuals(Rd @ R0, Rd @ t0 + d[3:], e9, s7)
    # scipy's cost: 1/2 sum f_scale^2 rho((f / f_scale)^2), rho(z) = z (z <= 1), 2 sqrt(z) - 1 (z > 1) == ceres::HuberLoss(0.1) on s = r^2
    sol = least_squares(f, np.zeros(6), loss="huber", f_scale=0.1, xtol=1e-15, ftol=1e-15, gtol=1e-15, jac="3-point", max_nfev=400)
    Rd = rodrigues(sol.x[:3])
    return Rd @ R0, Rd @ t0 + sol.x[3:], sol.cost


def lm_fixed_point(pfb, O, frames=range(0, 50, 5)):
    rows = []
    rng = np.random.default_rng(11)
    for k in frames:
        maps, qs, pose = frame_pair(pfb, O, k)
        start = pose.copy()
        start[4:] += rng.normal(0, 0.05, 3)                       # start away from the answer
        e9, s7 = residual_set(O, maps, qs, start)
        x, it, cost = O.lm_solve_ex(start, e9, s7, max_iter=200, ftol=1e-15)
        Rs, ts, cs = scipy_fixed_point(e9, s7, start)
        x4, it4, cost4 = O.lm_solve(start, e9, s7)                # the reference's cap: 4 iterations, ftol 1e-6
        rows.append({"frame": int(k), "edge_blocks": int(len(e9)), "surf_blocks": int(len(s7)), "oracle_iterations_uncapped": int(it),
                     "max_abs_R_diff": float(np.abs(quat_to_R(x[:4]) - Rs).max()), "max_abs_t_diff_m": float(np.abs(x[4:] - ts).max()),
                     "rel_cost_diff": float(abs(cost - cs) / cs),
                     "capped_vs_converged_t_diff_m": float(np.abs(x4[4:] - x[4:]).max()), "capped_iterations": int(it4)})
    return rows


# ---------------------------------------------------------------------------------------------------------------------------
# (ii) threshold flips
# ---------------------------------------------------------------------------------------------------------------------------
def numpy_fit_decisions(kind, map_pts, idx):
    """Per query with a valid 5-NN: does the reference's geometric test pass?  LAPACK eigh / SVD instead of Jacobi / QR."""
    nb = np.stack([map_pts["x"][idx], map_pts["y"][idx], map_pts["z"][idx]], 2).astype(np.float64)      # (n, 5, 3)
    if kind == 0:
        c = nb.sum(1) / 5.0
        d = nb - c[:, None, :]
        cov = np.einsum("nki,nkj->nij", d, d)
        w = np.linalg.eigvalsh(cov)
        return w[:, 2] > 3 * w[:, 1], w[:, 2] - 3 * w[:, 1], w[:, 2]
    x = np.einsum("nij,nj->ni", np.linalg.pinv(nb), -np.ones((len(nb), 5)))
    nrm = np.linalg.norm(x, axis=1)
    n = x / nrm[:, None]
    dist = np.abs(np.einsum("nkj,nj->nk", nb, n) + (1.0 / nrm)[:, None])
    return (dist <= 0.2).all(1), 0.2 - dist.max(1), np.full(len(nb), 0.2)


def flip_report(pfb, O, capi=None, frames=range(0, 100), cfg="cfg2", params=(0, 0.4, 75)):
    """Runs the oracle odometry over the frames; at every frame the association pass of the stage tap is evaluated on the frame's
    down-sampled queries against the map BEFORE the update at the frame's final pose.  Counts decision flips oracle vs numpy (and
    oracle vs GPU when capi is given)."""
    p = pfb.synth.config(cfg)
    od = O.Odom(0.4, *params)
    tot = {k: {"queries": 0, "knn_valid": 0, "fit_ok_oracle": 0, "flips_oracle_vs_numpy": 0, "flips_gpu_vs_oracle": 0, "flag_diff_gpu_vs_oracle": 0,
               "geom_max_abs_diff_gpu_vs_oracle": 0.0, "min_abs_margin_of_flips": None, "near_threshold_1e-9": 0} for k in ("edge", "surf")}
    last = max(frames)
    for f in range(last + 1):
        s = pfb.synth.scan(p, f)
        r = O.extract(s, num_lines=p.sensor_lines, order=1)
        e, u = s[r["edge_idx"]], s[r["surf_idx"]]
        if f == 0:
            od.init_map(e, u)
            continue
        maps = [od.get_map(0), od.get_map(1)]
        pose = od.update(e, u)
        if f not in frames:
            continue
        qs = [O.voxel_downsample(to_points(e), 0.4), O.voxel_downsample(to_points(u), 0.8)]
        for kind, name in ((0, "edge"), (1, "surf")):
            T = tot[name]
            idx, d2 = O.knn5(maps[kind], transform_queries(pose, qs[kind]))
            valid = idx[:, 4] >= 0
            _, _, flag, geom = O.associate(kind, maps[kind], qs[kind], pose, *params)
            ok_np, margin, scale = numpy_fit_decisions(kind, maps[kind], idx[valid])
            ok_or = flag[valid] != 0
            flips = ok_np != ok_or
            T["queries"] += int(len(flag)); T["knn_valid"] += int(valid.sum()); T["fit_ok_oracle"] += int(ok_or.sum())
            T["flips_oracle_vs_numpy"] += int(flips.sum())
            T["near_threshold_1e-9"] += int((np.abs(margin) <= 1e-9 * scale).sum())
            if flips.any():
                m = float(np.abs(margin[flips]).min())
                T["min_abs_margin_of_flips"] = m if T["min_abs_margin_of_flips"] is None else min(m, T["min_abs_margin_of_flips"])
            if capi is not None:
                _, _, gflag, ggeom = capi.associate(kind, maps[kind], qs[kind], pose, *params)
                T["flips_gpu_vs_oracle"] += int(((gflag != 0) != (flag != 0)).sum())
                T["flag_diff_gpu_vs_oracle"] += int((gflag != flag).sum())
                both = (gflag == 2) & (flag == 2)
                if both.any():
                    T["geom_max_abs_diff_gpu_vs_oracle"] = max(T["geom_max_abs_diff_gpu_vs_oracle"], float(np.abs(ggeom[both] - geom[both]).max()))
    return tot


# ---------------------------------------------------------------------------------------------------------------------------
# (iii) noise floor of the open conventions
# ---------------------------------------------------------------------------------------------------------------------------
def ate(poses, gt):
    rel = gt[:, 4:] - gt[0, 4:]
    return float(np.sqrt(((poses[:, 4:] - rel) ** 2).sum(1).mean()))


def oracle_trajectory(pfb, O, scans, num_lines, params, surf_order, literal_sort):
    """surf_order 1: ring position (ours), 0: ascending curvature (the reference, via its own compiled source when present)."""
    old = O.set_sort_mode(literal_sort)
    try:
        od = O.Odom(0.4, *params)
        poses = []
        for k, s in enumerate(scans):
            if surf_order == 0 and O.have_ref():
                ei, ui = O.ref_extract(s, num_lines=num_lines)
            else:
                r = O.extract(s, num_lines=num_lines, order=surf_order)
                ei, ui = r["edge_idx"], r["surf_idx"]
            if k == 0:
                od.init_map(s[ei], s[ui])
                poses.append(np.array([0, 0, 0, 1, 0, 0, 0.0]))
            else:
                poses.append(od.update(s[ei], s[ui]))
        return np.array(poses), od.stats()
    finally:
        O.set_sort_mode(old)


def noise_floor(pfb, O, capi=None, nframes=100, cfg="cfg2", params=(0, 0.4, 75)):
    p = pfb.synth.config(cfg)
    scans = [pfb.synth.scan(p, f) for f in range(nframes)]
    gt = np.array([pfb.synth.pose(p, f) for f in range(nframes)])
    runs = {}
    for name, so, lit in (("ours_convention(ring-position surf order, stable voxel sort)", 1, 0), ("reference_surf_order", 0, 0),
                          ("reference_voxel_sort", 1, 1), ("reference_both(closest to the real reference)", 0, 1)):
        poses, st = oracle_trajectory(pfb, O, scans, p.sensor_lines, params, so, lit)
        runs[name] = poses
    base = runs["ours_convention(ring-position surf order, stable voxel sort)"]
    out = {"frames": nframes, "config": cfg, "params": list(params), "ate_m": {k: ate(v, gt) for k, v in runs.items()},
           "max_abs_translation_diff_vs_ours_convention_m": {k: float(np.abs(v[:, 4:] - base[:, 4:]).max()) for k, v in runs.items()}}
    a0 = out["ate_m"]["ours_convention(ring-position surf order, stable voxel sort)"]
    out["ate_spread_rel"] = {k: abs(v - a0) / a0 for k, v in out["ate_m"].items()}
    if capi is not None:
        ex = capi.Extractor(num_lines=p.sensor_lines, max_points=131072)
        od = capi.Odometry(0.4, *params, max_map_points=1 << 19)
        gp = np.array([capi.frame_process(ex, od, s) for s in scans])
        ex.close(); od.close()
        out["ate_m"]["gpu"] = ate(gp, gt)
        out["ate_spread_rel"]["gpu"] = abs(out["ate_m"]["gpu"] - a0) / a0
        out["max_abs_translation_diff_vs_ours_convention_m"]["gpu"] = float(np.abs(gp[:, 4:] - base[:, 4:]).max())
        out["gpu_vs_reference_both"] = {"ate_rel_diff": abs(out["ate_m"]["gpu"] - out["ate_m"]["reference_both(closest to the real reference)"]) /
                                        out["ate_m"]["reference_both(closest to the real reference)"],
                                        "max_abs_translation_diff_m": float(np.abs(gp[:, 4:] - runs["reference_both(closest to the real reference)"][:, 4:]).max())}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=100)
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from pf_loader import pfb
    import oracle as O
    capi = None
    if a.gpu:
        capi = pfb.capi
        capi.lib()
    res = {"lm_fixed_point_vs_scipy": lm_fixed_point(pfb, O),
           "threshold_flips": flip_report(pfb, O, capi, frames=range(0, a.frames)),
           "noise_floor": noise_floor(pfb, O, capi, nframes=a.frames)}
    txt = json.dumps(res, indent=1)
    print(txt)
    if a.out:
        open(os.path.join(ROOT, a.out), "w").write(txt + "\n")


if __name__ == "__main__":
    main()
