"""GPU parity of the odometry stages against the oracle restatement, through the C-ABI stage taps."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _features(pfb, oracle, frame):
    p = pfb.synth.config("cfg2")
    s = pfb.synth.scan(p, frame)
    r = oracle.extract(s, order=1)
    return s[r["edge_idx"]], s[r["surf_idx"]]


def _maps_and_queries(pfb, oracle, capi):
    """A realistic (map, query) pair: frame-0 features voxelised as map, frame-1 features down-sampled as queries."""
    e0, s0 = _features(pfb, oracle, 0)
    e1, s1 = _features(pfb, oracle, 1)
    map_e = oracle.map_update(capi.make_points(e0[:, :3]), (0, 0, 0), 0.4, 0, 0.0, 0)
    map_s = oracle.map_update(capi.make_points(s0[:, :3]), (0, 0, 0), 0.8, 0, 0.0, 0)
    q_e = oracle.voxel_downsample(capi.make_points(e1[:, :3]), 0.4)
    q_s = oracle.voxel_downsample(capi.make_points(s1[:, :3]), 0.8)
    return (map_e, q_e), (map_s, q_s), (capi.make_points(e0[:, :3]), capi.make_points(s0[:, :3]))


POSE1 = np.array([0.0, 0.0, 0.0031, 0.999995, 1.0, 0.094, 0.0])


def test_knn5_matches_bruteforce_and_kdtree(capi, oracle, pfb):
    (map_e, q_e), (map_s, q_s), (raw_e, raw_s) = _maps_and_queries(pfb, oracle, capi)
    for m, q in ((map_e, q_e), (map_s, q_s), (raw_s, q_s)):     # raw_s: un-voxelised first-frame map (many points per cell)
        q4 = np.zeros((len(q), 4), np.float32)
        q4[:, 0], q4[:, 1], q4[:, 2] = q["x"] + 1.0, q["y"] + 0.09, q["z"]
        idx, d2 = capi.knn5(m, q4)
        ridx, rd2 = oracle.knn5(m, q4, mode=1)
        assert np.array_equal(idx, ridx)
        assert np.array_equal(d2.view(np.uint32), rd2.view(np.uint32))     # bit-exact float distances
        sub = slice(0, 300)
        bidx, bd2 = oracle.knn5(m, q4[sub], mode=0)
        assert np.array_equal(idx[sub], bidx) and np.array_equal(d2[sub].view(np.uint32), bd2.view(np.uint32))
        assert (idx[:, 0] >= 0).mean() > 0.3


def test_knn5_ties_and_sparse(capi, oracle):
    # regular lattice: many exactly equal distances -> ties must resolve to the lower map index
    g = np.arange(-3, 4, dtype=np.float32) * 0.5
    xyz = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
    m = capi.make_points(xyz)
    q4 = np.zeros((64, 4), np.float32)
    q4[:, :3] = xyz[::5][:64] + np.float32(0.25)
    idx, d2 = capi.knn5(m, q4)
    ridx, rd2 = oracle.knn5(m, q4, mode=0)
    assert np.array_equal(idx, ridx) and np.array_equal(d2.view(np.uint32), rd2.view(np.uint32))
    # a query far from everything is reported invalid
    far = np.array([[50, 50, 50, 0]], np.float32)
    idx, d2 = capi.knn5(m, far)
    assert (idx == -1).all() and np.isinf(d2).all()
    # fewer than 5 map points
    idx, _ = capi.knn5(m[:3], q4[:4])
    assert (idx == -1).all()


@pytest.mark.parametrize("params", [(0, 0.4, 75), (0, 0.0, 0), (0, 1.0, 200)])
def test_associate_matches_oracle(capi, oracle, pfb, params):
    (map_e, q_e), (map_s, q_s), _ = _maps_and_queries(pfb, oracle, capi)
    k_new, theta_p, theta_max = params
    rng = np.random.default_rng(3)
    for kind, (m, q) in ((0, (map_e, q_e)), (1, (map_s, q_s))):
        m = m.copy()
        m["r"] = rng.integers(0, 40, len(m))       # exercise the persistence rule
        m["g"] = rng.integers(0, 256, len(m))
        gm, gq, gflag, ggeom = capi.associate(kind, m, q, POSE1, k_new, theta_p, theta_max)
        om, oq, oflag, ogeom = oracle.associate(kind, m, q, POSE1, k_new, theta_p, theta_max)
        assert (oflag > 0).sum() > 100
        assert np.array_equal(gflag, oflag)                       # geometric validity + PFilter skip decisions
        assert gm.tobytes() == om.tobytes()                       # map observe counters after the sequential pass
        assert gq.tobytes() == oq.tobytes()                       # query r / g
        v = oflag > 0
        if kind == 1:
            np.testing.assert_allclose(ggeom[v, :4], ogeom[v, :4], rtol=1e-9, atol=1e-11)
        else:
            # the line direction sign is arbitrary: a and b may be swapped
            a_g, b_g, a_o, b_o = ggeom[v, :3], ggeom[v, 3:6], ogeom[v, :3], ogeom[v, 3:6]
            same = np.abs(a_g - a_o).max(1) < 1e-7
            swap = np.abs(a_g - b_o).max(1) < 1e-7
            assert (same | swap).all()
            np.testing.assert_allclose(a_g + b_g, a_o + b_o, rtol=1e-9, atol=1e-10)


def _residual_arrays(capi, oracle, pfb):
    (map_e, q_e), (map_s, q_s), _ = _maps_and_queries(pfb, oracle, capi)
    _, qe, fe, ge = oracle.associate(0, map_e, q_e, POSE1, 0, 0.0, 0)
    _, qs, fs, gs = oracle.associate(1, map_s, q_s, POSE1, 0, 0.0, 0)
    ve, vs = fe == 2, fs == 2
    pe = np.stack([q_e["x"], q_e["y"], q_e["z"]], 1).astype(np.float64)[ve]
    ps = np.stack([q_s["x"], q_s["y"], q_s["z"]], 1).astype(np.float64)[vs]
    edge9 = np.concatenate([pe, ge[ve, :6]], 1)
    surf7 = np.concatenate([ps, gs[vs, :4]], 1)
    return edge9, surf7


def test_eval_normal_equations(capi, oracle, pfb):
    edge9, surf7 = _residual_arrays(capi, oracle, pfb)
    for pose in (POSE1, np.array([0.01, -0.02, 0.03, 0.9993, 1.2, 0.0, -0.1])):
        H, g, c = capi.eval_normal_eq(pose, edge9, surf7)
        Ho, go, co = oracle.eval_normal_eq(pose, edge9, surf7)
        np.testing.assert_allclose(H, Ho, rtol=1e-10, atol=1e-9)
        np.testing.assert_allclose(g, go, rtol=1e-9, atol=1e-9)
        assert abs(c - co) <= 1e-12 * max(1.0, abs(co))
    H, g, c = capi.eval_normal_eq(POSE1, edge9[:0], surf7[:0])
    assert c == 0 and not H.any()


def test_lm_solve_matches_oracle(capi, oracle, pfb):
    edge9, surf7 = _residual_arrays(capi, oracle, pfb)
    start = POSE1 + np.array([0, 0, 0.002, 0, 0.05, -0.03, 0.01])
    x, it, cost = capi.lm_solve(start, edge9, surf7)
    xo, ito, costo = oracle.lm_solve(start, edge9, surf7)
    assert it == ito
    np.testing.assert_allclose(x, xo, rtol=1e-4, atol=1e-7)      # tolerance stated by the north star: 1e-4 relative
    assert abs(cost - costo) <= 1e-6 * costo
    assert np.abs(x[4:] - POSE1[4:]).max() < 0.02                # and it actually converged near the true pose
