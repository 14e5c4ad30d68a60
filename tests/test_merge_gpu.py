"""GPU parity of the streaming map update (pf_map_merge, csrc/merge.cu) against the oracle's full re-voxelisation
(rgbds + CropBox + extractstablepoint + r update, /root/reference/src/odomEstimationClass.cpp:606-647).

The merge consumes a map whose first part is already sorted by voxel key (what the previous update left) plus unsorted
extra points, and must produce the same voxels, centroids (bit-exact: canonical summation order) and counters as
re-sorting everything; only the position of "exceptions" (centroids that float rounding pushed out of their voxel) differs:
they trail the sorted part.  So maps are compared as multisets, and the sorted part must be strictly ascending in key."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _cloud(capi, rng, n, extent, counters=True):
    xyz = (rng.random((n, 3), dtype=np.float32) - 0.5) * np.array(extent, np.float32)
    if counters:
        return capi.make_points(xyz, r=rng.integers(0, 256, n), g=rng.integers(0, 256, n), b=0, a=255)
    return capi.make_points(xyz)


def _canon(a):
    v = np.ascontiguousarray(a).view(np.uint8).reshape(len(a), 16)
    order = np.lexsort(v.T[::-1])
    return a[order]


def _voxel_key(p, leaf):
    leaf = np.float32(leaf)
    k = [np.floor(p[c] / leaf).astype(np.int64) for c in ("z", "y", "x")]
    return (k[0] + (1 << 20)) << 42 | (k[1] + (1 << 20)) << 21 | (k[2] + (1 << 20))


def _split_sorted(m, leaf):
    """(strictly key-ascending part, rest) of an oracle map: a centroid that float rounding pushed out of its voxel (likely when
    points sit exactly on voxel faces) breaks the order and must travel with the unsorted input, as pf_odom_update does."""
    import bisect
    keys = _voxel_key(m, leaf)
    tails, tail_idx, prev = [], [], np.full(len(m), -1)
    for i, k in enumerate(keys):                       # longest strictly increasing subsequence
        j = bisect.bisect_left(tails, k)
        if j == len(tails):
            tails.append(k); tail_idx.append(i)
        else:
            tails[j] = k; tail_idx[j] = i
        prev[i] = tail_idx[j - 1] if j else -1
    keep = np.zeros(len(m), bool)
    i = tail_idx[-1] if tail_idx else -1
    while i >= 0:
        keep[i] = True
        i = prev[i]
    return m[keep], m[~keep]


def _check(capi, oracle, sorted_map, extra, center, leaf, prm):
    out, ns = capi.map_merge(sorted_map, extra, center, leaf, *prm)
    ref = oracle.map_update(np.concatenate([sorted_map, extra]), center, leaf, *prm)
    assert len(out) == len(ref), (len(out), len(ref))
    assert _canon(out).tobytes() == _canon(ref).tobytes()
    keys = _voxel_key(out[:ns], leaf)
    assert np.all(np.diff(keys) > 0), "sorted part is not strictly ascending in voxel key"
    return out, ns


@pytest.mark.parametrize("prm", [(0, 0.4, 75), (0, 0.0, 0), (3, 0.6, 40)])
@pytest.mark.parametrize("leaf", [0.4, 0.8])
def test_merge_matches_full_revoxelisation(capi, oracle, prm, leaf):
    rng = np.random.default_rng(int(leaf * 10) + prm[2])
    center0 = (1.0, -2.0, 0.5)
    raw = _cloud(capi, rng, 120000, (230, 210, 12))
    m0 = oracle.map_update(raw, center0, leaf, *prm)          # a sorted map, one point per voxel
    extra = _cloud(capi, rng, 9000, (120, 100, 10))
    extra["r"] = 0
    center1 = (2.5, -1.0, 0.6)                                # the crop box moved: part of the old map falls out
    _check(capi, oracle, m0, extra, center1, leaf, prm)


def test_merge_from_unsorted_start(capi, oracle):
    """n_sorted = 0 (first update after initMapWithPoints): everything is unsorted input."""
    rng = np.random.default_rng(3)
    raw = _cloud(capi, rng, 90000, (150, 150, 10), counters=False)
    out, ns = _check(capi, oracle, raw[:0], raw, (0, 0, 0), 0.8, (0, 0.4, 75))
    assert ns > 0


def test_merge_chain_of_updates(capi, oracle):
    """Several updates in a row, feeding sorted part + exceptions back in like pf_odom_update does."""
    rng = np.random.default_rng(11)
    leaf, prm = 0.4, (0, 0.4, 75)
    cur = _cloud(capi, rng, 60000, (90, 90, 6), counters=False)
    ns = 0
    ref = cur.copy()
    for f in range(6):
        center = (0.8 * f, 0.1 * f, 0.0)
        add = _cloud(capi, rng, 5000, (70, 70, 6))
        add["r"] = rng.integers(0, 20, len(add)); add["g"] = rng.integers(0, 255, len(add))
        if f > 0:
            add["x"] += np.float32(0.8 * f)
        nxt, ns2 = capi.map_merge(cur[:ns], np.concatenate([cur[ns:], add]), center, leaf, *prm)
        ref = oracle.map_update(np.concatenate([ref, add]), center, leaf, *prm)
        assert len(nxt) == len(ref)
        assert _canon(nxt).tobytes() == _canon(ref).tobytes()
        cur, ns = nxt, ns2
        assert len(cur) - ns < 64            # exceptions are rare


def test_merge_empty_inputs(capi):
    e = capi.make_points(np.zeros((0, 3), np.float32))
    out, ns = capi.map_merge(e, e, (0, 0, 0), 0.4, 0, 0.4, 75)
    assert len(out) == 0 and ns == 0
    far = capi.make_points(np.full((50, 3), 400.0, np.float32))
    out, ns = capi.map_merge(e, far, (0, 0, 0), 0.4, 0, 0.4, 75)
    assert len(out) == 0


def test_merge_large_map_streams(capi, oracle):
    """2M-voxel sorted map + 20k new points (the shape of a long low-speed sequence, BASELINE configs[3])."""
    rng = np.random.default_rng(21)
    raw = _cloud(capi, rng, 3000000, (198, 198, 30))
    m0 = oracle.map_update(raw, (0, 0, 0), 0.4, 0, 1.0, 200)
    add = _cloud(capi, rng, 20000, (100, 100, 10))
    _check(capi, oracle, m0, add, (0.2, 0.1, 0.0), 0.4, (0, 1.0, 200))


@pytest.mark.parametrize("seed", range(6))
def test_merge_adversarial_small_inputs(capi, oracle, seed):
    """Points on voxel faces, on the crop-box faces, duplicates, one crowded voxel, negative coordinates, tiny maps."""
    rng = np.random.default_rng(100 + seed)
    leaf = [0.4, 0.8][seed % 2]
    prm = [(0, 0.4, 75), (2, 0.9, 10), (0, 0.0, 0)][seed % 3]
    center = (rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-1, 1))
    lo = np.float32(center[0] - 100)
    n = int(rng.integers(1, 4000))
    grid = np.round(rng.uniform(-120, 120, (n, 3)) / leaf) * leaf                 # exactly on voxel faces (as far as floats allow)
    xyz = np.where(rng.random((n, 1)) < 0.5, grid, rng.uniform(-120, 120, (n, 3))).astype(np.float32)
    xyz[: n // 10] = xyz[0]                                                        # a crowded voxel of duplicates
    xyz[n // 10: n // 8, 0] = lo                                                   # on the lower crop face (kept: inclusive)
    xyz[n // 8: n // 6, 0] = np.nextafter(lo, np.float32(-1e9))                    # just outside
    raw = capi.make_points(xyz, r=rng.integers(0, 256, n), g=rng.integers(0, 256, n), b=0, a=255)
    m0 = oracle.map_update(raw, center, leaf, *prm)
    k = int(rng.integers(0, 3000))
    m0_xyz = np.stack([m0["x"], m0["y"], m0["z"]], axis=1) if len(m0) else np.zeros((1, 3), np.float32)
    exy = np.where(rng.random((k, 1)) < 0.5, m0_xyz[rng.integers(0, len(m0_xyz), k)],     # land exactly on existing centroids
                   rng.uniform(-110, 110, (k, 3))).astype(np.float32)
    extra = capi.make_points(exy, r=rng.integers(0, 8, k), g=rng.integers(0, 256, k), b=0, a=255)
    center1 = (center[0] + 0.7, center[1] - 0.2, center[2])
    m0s, exc = _split_sorted(m0, leaf)
    assert len(exc) <= 4
    _check(capi, oracle, m0s, np.concatenate([exc, extra]), center1, leaf, prm)


@pytest.mark.parametrize("seed", range(4))
def test_voxel_downsample_adversarial(capi, oracle, seed):
    rng = np.random.default_rng(200 + seed)
    leaf = [0.4, 0.8][seed % 2]
    n = int(rng.integers(1, 6000))
    grid = np.round(rng.uniform(-60, 60, (n, 3)) / leaf) * leaf
    xyz = np.where(rng.random((n, 1)) < 0.6, grid, rng.uniform(-60, 60, (n, 3))).astype(np.float32)
    xyz[: n // 5] = xyz[0]                                   # hundreds of duplicates in one voxel (long ordered sum)
    pts = capi.make_points(xyz)
    a, b = capi.voxel_downsample(pts, leaf), oracle.voxel_downsample(pts, leaf)
    assert len(a) == len(b) and a.tobytes() == b.tobytes()
