"""GPU parity of the voxel kernels (pf_voxel_downsample = PCL VoxelGrid, pf_map_update = CropBox + rgbds +
extractstablepoint + r update) against the oracle restatement: voxel assignment, output order, counters and
keep/remove masks bit-exact; centroids bit-exact too because the summation order is canonical (ascending index)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_cloud(capi, rng, n, extent=(60, 60, 8), counters=False):
    xyz = (rng.random((n, 3), dtype=np.float32) - 0.5) * np.array(extent, np.float32)
    if counters:
        return capi.make_points(xyz, r=rng.integers(0, 256, n), g=rng.integers(0, 256, n), b=rng.integers(0, 4, n), a=255)
    return capi.make_points(xyz)


def _same(a, b):
    assert len(a) == len(b), (len(a), len(b))
    assert a.tobytes() == b.tobytes()


@pytest.mark.parametrize("n,leaf", [(1, 0.4), (37, 0.4), (5000, 0.4), (70000, 0.8), (200000, 0.4)])
def test_voxel_downsample_random(capi, oracle, n, leaf):
    rng = np.random.default_rng(n)
    pts = _rand_cloud(capi, rng, n)
    _same(capi.voxel_downsample(pts, leaf), oracle.voxel_downsample(pts, leaf))


def test_voxel_downsample_rgba_mean_and_empty(capi, oracle):
    rng = np.random.default_rng(5)
    pts = _rand_cloud(capi, rng, 20000, extent=(10, 10, 2), counters=True)
    _same(capi.voxel_downsample(pts, 0.8), oracle.voxel_downsample(pts, 0.8))
    assert len(capi.voxel_downsample(pts[:0], 0.4)) == 0


def test_voxel_downsample_real_features(capi, oracle, cfg2_scans):
    _, scans = cfg2_scans
    r = oracle.extract(scans[2], order=1)
    for idx, leaf in ((r["edge_idx"], 0.4), (r["surf_idx"], 0.8)):
        pts = capi.make_points(scans[2][idx][:, :3])
        _same(capi.voxel_downsample(pts, leaf), oracle.voxel_downsample(pts, leaf))


def test_voxel_downsample_leaf_too_small(capi):
    xyz = np.array([[0, 0, 0], [3000, 3000, 3000]], np.float32)
    with pytest.raises(capi.PfError):
        capi.voxel_downsample(capi.make_points(xyz), 0.001)


@pytest.mark.parametrize("params", [(0, 0.4, 75), (0, 0.0, 0), (0, 1.0, 200), (3, 0.6, 40)])
@pytest.mark.parametrize("n", [1000, 150000])
def test_map_update_random(capi, oracle, n, params):
    rng = np.random.default_rng(n + params[2])
    pts = _rand_cloud(capi, rng, n, extent=(260, 230, 20), counters=True)   # part of the cloud lies outside the +-100 m crop box
    center = (7.3, -4.1, 0.6)
    k_new, theta_p, theta_max = params
    for leaf in (0.4, 0.8):
        _same(capi.map_update(pts, center, leaf, k_new, theta_p, theta_max), oracle.map_update(pts, center, leaf, k_new, theta_p, theta_max))


def test_map_update_idempotent_voxels(capi, oracle):
    """A second update of an already voxelised map keeps exactly one point per voxel (centroids stay in their voxel)."""
    rng = np.random.default_rng(9)
    pts = _rand_cloud(capi, rng, 50000, extent=(80, 80, 6))
    m1 = capi.map_update(pts, (0, 0, 0), 0.4, 0, 0.0, 0)
    m2 = capi.map_update(m1, (0, 0, 0), 0.4, 0, 0.0, 0)
    assert len(m1) == len(m2)
    assert np.array_equal(m1["x"], m2["x"]) and np.array_equal(m2["r"], np.minimum(255, m1["r"].astype(int) + 2))
    _same(m2, oracle.map_update(m1, (0, 0, 0), 0.4, 0, 0.0, 0))


def test_map_update_all_cropped(capi):
    xyz = np.full((100, 3), 500.0, np.float32)
    assert len(capi.map_update(capi.make_points(xyz), (0, 0, 0), 0.4, 0, 0.4, 75)) == 0
