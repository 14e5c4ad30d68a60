"""BASELINE.json configs[3] at its stated size: 32-ring VLP-32-shaped scans, low speed (0.15 m per frame), PFilter 0/1/200, 2000 frames.
Default sequence: cfg4s (the planes+poles street).  The campus loop (cfg4) is not usable at this length: the reference algorithm itself
(oracle and GPU alike) loses its yaw estimate there after ~540 frames and its constant-velocity prediction runs away.

Runs the whole sequence through the GPU frame pipeline (pf_frame_submit / pf_frame_wait from pinned host scans), records the
local-map sizes every 100 frames and the throughput, runs the CPU oracle over the first --oracle frames and reports pose / map
parity (incl. a count of differing voxels), and measures the streaming map update (K9) on the map the pipeline itself grew.
--cfg cfg4sd --params 0,0,0 is the map-growth stress: the same drive with volumetric scatter beside the lane, filter off.

usage: run_cfg4.py [--cfg cfg4] [--frames 2000] [--oracle 300] [--params 0,1,200] [--out profiles/x1_cfg4.json]
"""
import argparse
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pf_loader import pfb  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg4s")
    ap.add_argument("--frames", type=int, default=2000)
    ap.add_argument("--oracle", type=int, default=300)
    ap.add_argument("--params", default="0,1,200")
    ap.add_argument("--map-cap", type=int, default=1 << 22)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    k_new, theta_p, theta_max = args.params.split(",")
    prm = (int(k_new), float(theta_p), int(theta_max))
    capi = pfb.capi
    import oracle as O
    from parity_utils import map_diff
    p = pfb.synth.config(args.cfg)
    N = args.frames
    pfb.synth.scan(p, 0)     # builds the scene once; the generator is read-only afterwards
    t0 = time.perf_counter()
    with ThreadPoolExecutor(os.cpu_count() or 4) as tp:
        scans = list(tp.map(lambda f: pfb.synth.scan(p, f), range(N)))
    t_gen = time.perf_counter() - t0
    gt = np.array([pfb.synth.pose(p, f) for f in range(N)])
    pinned = []
    for s in scans:
        a, ptr = capi.pinned_array((len(s), 4), np.float32)
        a[:] = s
        pinned.append((a, ptr))

    def handles():
        return (capi.Extractor(num_lines=p.sensor_lines, max_points=p.sensor_lines * 1800),
                capi.Odometry(0.4, *prm, max_map_points=args.map_cap, max_features=p.sensor_lines * 1800))

    # warm-up on throw-away handles
    ex, od = handles()
    for k in range(4):
        capi.frame_process(ex, od, pinned[k][0])
    ex.close(); od.close()

    ex, od = handles()
    poses, sizes, seg_ms = [], [], []
    t_all = time.perf_counter()
    t_seg = t_all
    fid_prev = capi.frame_submit(ex, od, pinned[0][0])
    for k in range(1, N):
        fid = capi.frame_submit(ex, od, pinned[k][0])
        poses.append(capi.frame_wait(od, fid_prev))
        fid_prev = fid
        if k % 100 == 0:
            now = time.perf_counter()
            seg_ms.append(1e3 * (now - t_seg) / 100)
            st = od.stats()          # synchronises: once per 100 frames, outside the per-segment clock
            sizes.append({"frame": k, "map_edge": st["map_edge"], "map_surf": st["map_surf"], "n_edge_ds": st["n_edge_ds"], "n_surf_ds": st["n_surf_ds"]})
            t_seg = time.perf_counter()
    poses.append(capi.frame_wait(od, fid_prev))
    t_gpu = time.perf_counter() - t_all
    poses = np.array(poses)
    st = od.stats()
    sizes.append({"frame": N - 1, "map_edge": st["map_edge"], "map_surf": st["map_surf"], "n_edge_ds": st["n_edge_ds"], "n_surf_ds": st["n_surf_ds"]})
    gmaps = [od.map_part(0), od.map_part(1)]
    graph_captures = od.graph_captures
    last_pose = poses[-1]
    ex.close(); od.close()

    rel = gt[:, 4:] - gt[0, 4:]
    ate = float(np.sqrt(((poses[:, 4:] - rel) ** 2).sum(1).mean()))
    out = {"config": f"{args.cfg}: {p.sensor_lines}-ring, {N} frames, PFilter {args.params}, map_resolution 0.4",
           "points_per_scan": int(np.mean([len(s) for s in scans])), "scan_generation_s": t_gen,
           "gpu": {"scans_per_s_e2e": N / t_gpu, "ms_per_frame": 1e3 * t_gpu / N, "api": "pf_frame_submit + pf_frame_wait, pinned host scans",
                   "ms_per_frame_by_100": seg_ms, "graph_captures": graph_captures, "ate_vs_ground_truth_m": ate},
           "map_sizes": sizes, "map_points_final": int(len(gmaps[0]) + len(gmaps[1]))}

    # K9 on the map the pipeline grew: sorted part of the final maps + one frame's worth of new points around the last pose
    peak = 6533.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    rng = np.random.default_rng(7)
    k9 = {}
    for which, leaf in ((0, 0.4), (1, 0.8)):
        m = gmaps[which]
        if len(m) < 16:
            continue
        from parity_utils import voxel_keys
        # the sorted part = strictly ascending (z, y, x) voxel order; exceptions trail it -> re-sort through pf_map_update semantics
        ms = capi.map_update(m, last_pose[4:], leaf, 0, 0.0, 0)     # filter off: pure re-voxelisation, sorted, same points
        add = capi.make_points((rng.random((5000, 3), dtype=np.float32) - 0.5) * np.array([80, 80, 8], np.float32) + last_pose[4:].astype(np.float32), r=0, g=1)
        t = capi.map_merge_timed(ms, add, last_pose[4:], leaf, *prm, reps=6)
        b = 16.0 * (len(ms) + len(add)) + 16.0 * t["n_out"]
        k9["edge" if which == 0 else "surf"] = {"map_points": int(len(ms)), "new_points": 5000, "ms_stream": t["ms_stream"], "ms_whole_update": t["ms_total"],
                                                 "algorithmic_bytes": b, "frac_stream": b / (t["ms_stream"] * 1e-3) / 1e9 / peak,
                                                 "frac_whole_update": b / (t["ms_total"] * 1e-3) / 1e9 / peak}
    out["k9_on_grown_map"] = k9

    # oracle over the first frames
    M = min(args.oracle, N)
    if M > 1:
        ref = O.Odom(0.4, *prm)
        rposes = []
        # GPU maps at frame M-1 need a second GPU pass that stops there
        ex, od = handles()
        gp = [capi.frame_process(ex, od, pinned[k][0]) for k in range(M)]
        assert np.array(gp).tobytes() == poses[:M].tobytes(), "blocking and queued frames disagree"
        gm = [od.map_part(0), od.map_part(1)]
        gst = od.stats()
        ex.close(); od.close()
        t0 = time.perf_counter()
        for k in range(M):
            s = scans[k]
            r = O.extract(s, num_lines=p.sensor_lines, order=1)
            if k == 0:
                ref.init_map(s[r["edge_idx"]], s[r["surf_idx"]])
                rposes.append(np.array([0, 0, 0, 1, 0, 0, 0.0]))
            else:
                rposes.append(ref.update(s[r["edge_idx"]], s[r["surf_idx"]]))
        t_cpu = time.perf_counter() - t0
        rposes = np.array(rposes)
        gpa = np.array(gp)
        rm = [ref.get_map(0), ref.get_map(1)]
        rst = ref.stats()
        out["parity_vs_oracle"] = {
            "frames": M, "max_abs_translation_diff_m": float(np.abs(gpa[:, 4:] - rposes[:, 4:]).max()),
            "max_abs_quaternion_diff": float(np.abs(gpa[:, :4] - rposes[:, :4]).max()),
            "map_edge": [int(len(gm[0])), int(len(rm[0]))], "map_surf": [int(len(gm[1])), int(len(rm[1]))],
            "n_ds_equal": bool(gst["n_edge_ds"] == rst["n_edge_ds"] and gst["n_surf_ds"] == rst["n_surf_ds"]),
            "voxel_diff_edge": map_diff(gm[0], rm[0], 0.4), "voxel_diff_surf": map_diff(gm[1], rm[1], 0.8),
            "cpu_oracle_scans_per_s": M / t_cpu,
            "ate_gpu_m": float(np.sqrt(((gpa[:, 4:] - rel[:M]) ** 2).sum(1).mean())),
            "ate_oracle_m": float(np.sqrt(((rposes[:, 4:] - rel[:M]) ** 2).sum(1).mean()))}
    for _, ptr in pinned:
        capi.host_free(ptr)
    txt = json.dumps(out, indent=1)
    print(txt)
    if args.out:
        with open(os.path.join(ROOT, args.out), "w") as f:
            f.write(txt + "\n")


if __name__ == "__main__":
    main()
