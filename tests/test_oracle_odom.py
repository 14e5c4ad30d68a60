"""CPU: numerical kernels of the odometry oracle against independent references (numpy.linalg, finite differences,
brute force), its regression golden, and size-independent properties of the voxel / map operators.
The reference ships no fixture for this part: "parity unpinned" -- see oracle/oracle_odom.cpp."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "odom_oracle_golden.npz")


def test_eig3_against_numpy(oracle):
    rng = np.random.default_rng(0)
    for _ in range(200):
        P = rng.normal(size=(5, 3)) * rng.uniform(0.01, 3, size=3)
        C = (P - P.mean(0)).T @ (P - P.mean(0))
        w, V = oracle.eig3(C)
        wn, Vn = np.linalg.eigh(C)
        np.testing.assert_allclose(w, wn, rtol=1e-10, atol=1e-12 * abs(wn).max())
        assert abs(abs(V[:, 2] @ Vn[:, 2]) - 1) < 1e-8
        np.testing.assert_allclose(C @ V, V * w, atol=1e-10 * abs(wn).max())


def test_plane_least_squares_against_numpy(oracle):
    rng = np.random.default_rng(1)
    for _ in range(200):
        n = rng.normal(size=3); n /= np.linalg.norm(n)
        pts = rng.normal(size=(5, 3)) * 2
        pts -= np.outer(pts @ n - rng.uniform(1, 20), n)          # on a plane n.x = d
        pts += rng.normal(size=(5, 3)) * 0.01
        x = oracle.qr_solve(pts, -np.ones(5))
        xn = np.linalg.lstsq(pts, -np.ones(5), rcond=None)[0]
        np.testing.assert_allclose(x, xn, rtol=1e-7, atol=1e-9)


def test_analytic_jacobians_against_finite_differences(oracle):
    rng = np.random.default_rng(2)
    for kind in (0, 1):
        for _ in range(20):
            q = rng.normal(size=4); q /= np.linalg.norm(q)
            pose = np.concatenate([q, rng.normal(size=3)])
            p = rng.normal(size=3) * 5
            if kind == 0:
                a = rng.normal(size=3) * 5
                geom = np.concatenate([p, a, a + rng.normal(size=3)])
            else:
                n = rng.normal(size=3); n /= np.linalg.norm(n)
                geom = np.concatenate([p, n, [rng.normal()]])
            r0, J = oracle.residual(kind, geom, pose)
            for k in range(6):
                d = np.zeros(6); d[k] = 1e-6
                rp, _ = oracle.residual(kind, geom, oracle.se3_plus(pose, d))
                rm, _ = oracle.residual(kind, geom, oracle.se3_plus(pose, -d))
                assert abs((rp - rm) / 2e-6 - J[k]) < 1e-5 * max(1.0, abs(J[k]))


def test_se3_plus_is_left_multiplication(oracle):
    x = np.array([0, 0, 0, 1, 1.0, 2.0, 3.0])
    out = oracle.se3_plus(x, np.array([0, 0, np.pi / 2, 0, 0, 0]))
    np.testing.assert_allclose(out[:4], [0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)], atol=1e-12)
    np.testing.assert_allclose(out[4:], [-2.0, 1.0, 3.0], atol=1e-12)   # translation is rotated too (:92)


def test_kdtree_equals_bruteforce_including_ties(oracle):
    rng = np.random.default_rng(3)
    xyz = np.round(rng.uniform(-4, 4, size=(3000, 3)) * 4) / 4         # coarse lattice: many exact distance ties
    m = np.zeros(len(xyz), oracle.POINT_DTYPE)
    m["x"], m["y"], m["z"] = xyz.T
    q = np.zeros((400, 4), np.float32)
    q[:, :3] = rng.uniform(-4, 4, size=(400, 3))
    q[:100, :3] = np.round(q[:100, :3] * 2) / 2
    ia, da = oracle.knn5(m, q, mode=0)
    ib, db = oracle.knn5(m, q, mode=1)
    assert np.array_equal(ia, ib) and np.array_equal(da.view(np.uint32), db.view(np.uint32))
    ok = ia[:, 0] >= 0
    d = ((q[ok, None, :3].astype(np.float32) - xyz[ia[ok]].astype(np.float32)) ** 2)
    assert np.all(np.diff(da[ok], axis=1) >= 0) and ok.sum() > 50
    np.testing.assert_allclose(d.sum(-1), da[ok], rtol=1e-6)


def test_voxel_downsample_properties(oracle):
    rng = np.random.default_rng(4)
    pts = np.zeros(20000, oracle.POINT_DTYPE)
    xyz = (rng.random((20000, 3), dtype=np.float32) - 0.5) * np.array([40, 40, 6], np.float32)
    pts["x"], pts["y"], pts["z"] = xyz.T
    pts["a"] = 255
    leaf = np.float32(0.8)
    out = oracle.voxel_downsample(pts, float(leaf))
    inv = np.float32(1) / leaf
    key = np.floor(xyz * inv).astype(np.int64)
    uniq = np.unique(key, axis=0)
    assert len(out) == len(uniq)                                   # one output per occupied voxel
    okey = np.floor(np.stack([out["x"], out["y"], out["z"]], 1) * inv).astype(np.int64)
    lin = lambda k: (k[:, 2] - key[:, 2].min()) * 10**8 + (k[:, 1] - key[:, 1].min()) * 10**4 + (k[:, 0] - key[:, 0].min())
    assert np.all(np.diff(lin(okey)) > 0)                          # ascending (z, y, x) voxel order, centroid stays in its voxel
    assert abs(out["x"].astype(np.float64).mean() - xyz[:, 0].mean()) < 0.2
    again = oracle.voxel_downsample(out, float(leaf))
    assert again.tobytes() == out.tobytes()                        # idempotent


def test_map_update_rules(oracle):
    pts = np.zeros(6, oracle.POINT_DTYPE)
    pts["x"] = [0.1, 0.15, 5.0, 300.0, 9.0, 9.05]
    pts["r"] = [10, 4, 20, 0, 252, 3]
    pts["g"] = [3, 9, 2, 0, 1, 200]
    pts["a"] = 255
    out = oracle.map_update(pts, (0, 0, 0), 0.4, 0, 0.4, 75)
    # voxel 0: r=max(10,4)=10, g=max(3,9)=9 -> 9 >= 10*0.4 keep, r -> 12; x=5: g=2 < 8 & r>0 & g<76 -> deleted;
    # x=300 cropped; voxel at 9: r=252,g=200 -> g >= theta_max+1 keeps, r saturates to 255
    assert len(out) == 2
    assert (out["r"].tolist(), out["g"].tolist()) == ([12, 255], [9, 200])
    assert out["b"].tolist() == [0, 0] and out["a"].tolist() == [255, 255]
    off = oracle.map_update(pts, (0, 0, 0), 0.4, 0, 0.0, 0)       # F-LOAM mode: nothing is ever deleted
    assert len(off) == 3


def test_lm_recovers_a_known_pose(oracle):
    rng = np.random.default_rng(5)
    q = np.array([0.01, -0.02, 0.03, 0]); q[3] = np.sqrt(1 - (q ** 2).sum())
    true = np.concatenate([q, [0.3, -0.2, 0.1]])

    def apply(pose, p):
        qv, w = pose[:3], pose[3]
        uv = 2 * np.cross(qv, p)
        return p + w * uv + np.cross(qv, uv) + pose[4:]
    surf = []
    for n in np.eye(3):
        for _ in range(60):
            p = rng.uniform(-10, 10, size=3)
            surf.append(np.concatenate([p, n, [-(n @ apply(true, p))]]))
    edge = []
    for _ in range(60):
        p = rng.uniform(-10, 10, size=3)
        w = apply(true, p)
        d = rng.normal(size=3); d /= np.linalg.norm(d)
        edge.append(np.concatenate([p, w + 0.1 * d, w - 0.1 * d]))
    x = np.array([0, 0, 0, 1, 0, 0, 0.0])
    for _ in range(3):                                            # three ceres::Solve calls of <= 4 iterations each
        x, it, cost = oracle.lm_solve(x, np.array(edge), np.array(surf))
    np.testing.assert_allclose(x, true, atol=1e-6)
    assert cost < 1e-10


def test_regression_golden(oracle, pfb):
    g = np.load(GOLDEN)
    lines, az, seed = (int(v) for v in g["synth"])
    p = pfb.synth.params(sensor_lines=lines, azimuth_steps=az, seed=seed)
    od = oracle.Odom(0.4, 0, 0.4, 75)
    for f in range(6):
        s = pfb.synth.scan(p, f)
        r = oracle.extract(s, order=1)
        e, u = s[r["edge_idx"]], s[r["surf_idx"]]
        if f == 0:
            od.init_map(e, u)
            continue
        pose = od.update(e, u)
        st = od.stats()
        np.testing.assert_allclose(pose, g["poses"][f - 1], rtol=0, atol=1e-9)
        assert [st["n_edge_ds"], st["n_surf_ds"], st["n_edge_res"], st["n_surf_res"], st["map_edge"], st["map_surf"]] == g["sizes"][f - 1].tolist()
