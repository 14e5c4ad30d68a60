"""CPU: the oracle restatement of LaserMappingClass (oracle/oracle_mapping.cpp) against an independent numpy model of
/root/reference/src/laserMappingClass.cpp:152-208 on small inputs (one VoxelGrid pass per 50 m cell, cells in x, y, z order)."""
import numpy as np


def _numpy_model(pts, leaf):
    f = np.float32
    inv = f(1.0) / f(leaf)
    cell = np.floor(pts[:, :3].astype(np.float64) / 50.0 + 0.5).astype(np.int64)
    vox = np.floor(pts[:, :3] * inv).astype(np.int64)
    inten = np.minimum(1.0, np.maximum(pts[:, 2].astype(np.float64) + 2.0, 0.0) / 5).astype(f)
    key = {}
    for i in range(len(pts)):
        k = (cell[i, 0], cell[i, 1], cell[i, 2], vox[i, 2], vox[i, 1], vox[i, 0])
        key.setdefault(k, []).append(i)
    out = []
    for k in sorted(key):
        acc = np.zeros(4, f)
        for i in key[k]:
            acc = (acc + np.array([pts[i, 0], pts[i, 1], pts[i, 2], inten[i]], f)).astype(f)
        out.append(acc / f(len(key[k])))
    return np.array(out, f).reshape(-1, 4)


def test_single_update_identity_pose(oracle):
    rng = np.random.default_rng(1)
    pts = ((rng.random((4000, 4), dtype=np.float32) - 0.5) * np.array([180, 160, 12, 1], np.float32)).astype(np.float32)
    om = oracle.Mapping(0.4)
    om.update(pts, np.eye(4)[:3].reshape(12))
    got = om.get_map()
    want = _numpy_model(pts, 0.4)
    assert om.dropped() == 0
    assert got.shape == want.shape and got.tobytes() == want.tobytes()


def test_second_update_is_incremental(oracle):
    """Re-filtering an already filtered cell keeps its centroids; new points merge with weight 1 per old centroid."""
    rng = np.random.default_rng(2)
    a = ((rng.random((3000, 4), dtype=np.float32) - 0.5) * np.array([60, 60, 6, 1], np.float32)).astype(np.float32)
    om = oracle.Mapping(0.8)
    rt = np.eye(4)[:3].reshape(12)
    om.update(a, rt)
    m1 = om.get_map()
    om.update(a[:0], rt)
    assert om.get_map().tobytes() == m1.tobytes()
    far = np.array([[400.0, 0, 0, 0]], np.float32)
    om.update(far, rt)
    assert om.dropped() == 1 and len(om.get_map()) == len(m1)
