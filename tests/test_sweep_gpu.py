"""Map-size sweep kernels (BASELINE.json configs[4]): exact 5-NN and the grid-wide normal-equation kernel against the CPU oracle
at sweep sizes the oracle finishes in seconds, and size-independent properties at the large sizes."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m_points", [1_000_000, 2_000_000])
def test_knn5_matches_kdtree_at_sweep_sizes(capi, oracle, m_points):
    import sweep
    xyz = sweep.plane_map(m_points)[:m_points]
    pts = capi.make_points(xyz, r=0, g=1)
    rng = np.random.default_rng(m_points)
    nq = 20000
    q = np.zeros((nq, 4), np.float32)
    q[:, :3] = xyz[rng.integers(0, m_points, nq)] + rng.normal(0, 0.2, (nq, 3)).astype(np.float32)
    q[:50, :3] += 500.0                                   # far outside the map: no neighbour within 1 m
    gi, gd = capi.knn5(pts, q)
    ci, cd, _, _ = oracle.knn5_timed(pts, q)
    assert np.array_equal(gi, ci)                         # indices bit-exact, ties by lower index
    assert gd.tobytes() == cd.tobytes()                   # float L2_Simple distances bit-exact
    assert (gi[:50] == -1).all() and (gi[50:, 4] >= 0).mean() > 0.99


def test_normal_eq_stream_matches_oracle(capi, oracle):
    import sweep
    e, s = sweep.residual_blocks(150_001)                 # ragged last tiles on both kinds
    s = s[:140_003]
    H, g, cost, ms = capi.eval_normal_eq_timed(sweep.SWEEP_POSE, e, s, reps=1)
    Ho, go, co = oracle.eval_normal_eq(sweep.SWEEP_POSE, e, s)
    np.testing.assert_allclose(H, Ho, rtol=1e-10, atol=1e-8)
    np.testing.assert_allclose(g, go, rtol=1e-10, atol=1e-8)
    assert abs(cost - co) <= 1e-10 * abs(co)
    # the per-frame cluster kernel gives the same sums on the same blocks
    Hc, gc, cc = capi.eval_normal_eq(sweep.SWEEP_POSE, e, s)
    np.testing.assert_allclose(H, Hc, rtol=1e-11, atol=1e-9)
    # above 262144 blocks pf_eval_normal_eq itself routes to the grid-wide kernel
    e2, s2 = sweep.residual_blocks(400_000)
    H2, g2, c2 = capi.eval_normal_eq(sweep.SWEEP_POSE, e2, s2)
    H3, g3, c3, _ = capi.eval_normal_eq_timed(sweep.SWEEP_POSE, e2, s2, reps=1)
    assert H2.tobytes() == H3.tobytes() and c2 == c3


def test_normal_eq_stream_edge_cases(capi, oracle):
    import sweep
    e, s = sweep.residual_blocks(1000)
    z9, z7 = np.zeros((0, 9)), np.zeros((0, 7))
    H, g, cost, _ = capi.eval_normal_eq_timed(sweep.SWEEP_POSE, z9, z7, reps=1)
    assert not H.any() and not g.any() and cost == 0.0
    for ee, ss in ((e, z7), (z9, s), (e[:1], s[:255]), (e[:256], s[:257])):
        H, g, cost, _ = capi.eval_normal_eq_timed(sweep.SWEEP_POSE, ee, ss, reps=1)
        Ho, go, co = oracle.eval_normal_eq(sweep.SWEEP_POSE, ee, ss)
        np.testing.assert_allclose(H, Ho, rtol=1e-11, atol=1e-10)
        np.testing.assert_allclose(g, go, rtol=1e-11, atol=1e-10)


def test_normal_eq_stream_linearity_at_full_size(capi):
    """Size-independent property at a sweep size the oracle does not reach: the sums over k copies of a block set are k times the
    sums of one copy (to summation rounding), and repeated launches are bit-identical (fixed reduction order)."""
    import sweep
    e1, s1 = sweep.residual_blocks(1 << 20)
    H1, g1, c1, _ = capi.eval_normal_eq_timed(sweep.SWEEP_POSE, e1, s1, reps=1)
    k = 5
    e5, s5 = np.tile(e1, (k, 1)), np.tile(s1, (k, 1))     # 10.5 M residual blocks
    H5, g5, c5, ms = capi.eval_normal_eq_timed(sweep.SWEEP_POSE, e5, s5, reps=2)
    np.testing.assert_allclose(H5, k * H1, rtol=1e-11)
    np.testing.assert_allclose(g5, k * g1, rtol=1e-9, atol=1e-6)
    assert abs(c5 - k * c1) <= 1e-11 * c5
    H5b, g5b, c5b, _ = capi.eval_normal_eq_timed(sweep.SWEEP_POSE, e5, s5, reps=1)
    assert H5.tobytes() == H5b.tobytes() and g5.tobytes() == g5b.tobytes() and c5 == c5b


@pytest.mark.parametrize("kind", ["dense", "lattice", "faces"])
def test_knn5_row_pruning_is_exact(capi, oracle, kind):
    """The search skips rows of cells whose lower distance bound exceeds the running fifth distance (knn.cuh).  Dense maps make it
    skip most; lattice maps make every distance tie; queries ON cell faces make the bounds zero -- results must stay FLANN's."""
    rng = np.random.default_rng({"dense": 1, "lattice": 2, "faces": 3}[kind])
    n, nq = 300_000, 30_000
    xyz = (rng.random((n, 3), dtype=np.float32) - 0.5) * np.array([40, 40, 10], np.float32)        # ~19 points per 1 m cell
    if kind == "lattice":
        xyz = np.round(xyz * 4) / 4                                                              # 0.25 m lattice: masses of exact ties
    q = np.zeros((nq, 4), np.float32)
    q[:, :3] = xyz[rng.integers(0, n, nq)] + rng.normal(0, 0.15, (nq, 3)).astype(np.float32)
    if kind == "faces":
        c = int(rng.integers(0, 3))
        q[: nq // 2, c] = np.round(q[: nq // 2, c])                                              # integer coordinate: on a cell face
        q[nq // 2:, :3] = np.round(q[nq // 2:, :3])
    if kind == "lattice":
        q[:, :3] = np.round(q[:, :3] * 8) / 8
    pts = capi.make_points(xyz.astype(np.float32))
    gi, gd = capi.knn5(pts, q)
    ci, cd, _, _ = oracle.knn5_timed(pts, q)
    assert np.array_equal(gi, ci)
    assert gd.tobytes() == cd.tobytes()
    assert (gi[:, 4] >= 0).mean() > 0.95
