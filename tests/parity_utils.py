"""Helpers shared by the parity tests and the tools that report parity (test infrastructure; imports nothing from the product)."""
import numpy as np


def voxel_keys(pts, leaf):
    """rgbds voxel coordinates of map points: floor(x / leaf) in float (/root/reference/src/odomEstimationClass.cpp:63-68)."""
    leaf = np.float32(leaf)
    k = np.stack([np.floor(pts[c] / leaf).astype(np.int64) for c in ("x", "y", "z")], 1)
    return (k[:, 0] + 4096) | ((k[:, 1] + 4096) << 20) | ((k[:, 2] + 4096) << 40)


def map_diff(gm, rm, leaf, tol=1e-4):
    """Voxel-level difference of two local maps (structured arrays of pf_point).  Returns counts:
    only_a / only_b: occupied voxels present in one map only; counters: common voxels whose (r, g) differ;
    moved: common voxels whose centroid differs by more than tol in any coordinate; common: voxels in both."""
    ka, kb = voxel_keys(gm, leaf), voxel_keys(rm, leaf)
    ua, ia = np.unique(ka, return_index=True)
    ub, ib = np.unique(kb, return_index=True)
    common, ca, cb = np.intersect1d(ua, ub, return_indices=True)
    a, b = gm[ia[ca]], rm[ib[cb]]
    counters = int(((a["r"] != b["r"]) | (a["g"] != b["g"])).sum())
    moved = int(((np.abs(a["x"] - b["x"]) > tol) | (np.abs(a["y"] - b["y"]) > tol) | (np.abs(a["z"] - b["z"]) > tol)).sum())
    return {"only_a": int(len(ua) - len(common)), "only_b": int(len(ub) - len(common)), "common": int(len(common)),
            "counters": counters, "moved": moved, "dup_a": int(len(ka) - len(ua)), "dup_b": int(len(kb) - len(ub))}
