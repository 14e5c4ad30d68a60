"""GPU parity of the global map (pf_mapping_*, csrc/mapping.cu) against the oracle restatement of LaserMappingClass
(/root/reference/src/laserMappingClass.cpp:152-208): same cells, same voxels, bit-identical centroids / intensities; the
order equals the reference's getMap order except for the rare centroids that left their voxel (they trail the array)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _canon(a):
    v = np.ascontiguousarray(a).view(np.uint8).reshape(len(a), 16)
    return a[np.lexsort(v.T[::-1])]


def _rt(pfb, pose7):
    return pfb.capi.pose_to_rt(pose7)


def _feed(pfb, capi, oracle, cfg, frames, leaf=0.4, stride=1):
    p = pfb.synth.config(cfg)
    gm = capi.Mapping(leaf, max_map_points=4 << 20, max_points=131072)
    om = oracle.Mapping(leaf)
    for f in frames:
        s = pfb.synth.scan(p, f)[::stride]
        rt = _rt(pfb, pfb.synth.pose(p, f))
        gm.update(s, rt)
        om.update(s, rt)
    return gm, om


def test_mapping_sequence_matches_oracle(pfb, capi, oracle):
    gm, om = _feed(pfb, capi, oracle, "cfg2", range(0, 12))
    g, r = gm.get_map(), om.get_map()
    st = gm.stats()
    assert len(g) == len(r) and len(g) > 20000
    assert _canon(g).tobytes() == _canon(r).tobytes()
    ns = st["n_sorted"]
    assert len(g) - ns < 64
    if ns == len(g):
        assert g.tobytes() == r.tobytes()          # same order as the reference's getMap
    assert st["dropped"] == om.dropped()


def test_mapping_crosses_cell_boundaries(pfb, capi, oracle):
    """A trajectory that crosses 50 m cell faces: the 5x5x5 block moves, old cells pass through untouched."""
    p = pfb.synth.config("cfg2")
    gm = capi.Mapping(0.8, max_map_points=2 << 20, max_points=131072)
    om = oracle.Mapping(0.8)
    rng = np.random.default_rng(4)
    s0 = pfb.synth.scan(p, 0)[::3]
    for k in range(10):
        pose = np.array([0, 0, np.sin(0.05 * k), np.cos(0.05 * k), 23.0 * k, -9.0 * k, 0.3 * k])
        s = s0 + np.float32(0.01) * rng.standard_normal(s0.shape).astype(np.float32)
        rt = _rt(pfb, pose)
        gm.update(s, rt)
        om.update(s, rt)
    g, r = gm.get_map(), om.get_map()
    assert len(g) == len(r)
    assert _canon(g).tobytes() == _canon(r).tobytes()
    assert gm.stats()["dropped"] == om.dropped()


def test_mapping_class_interface_and_empty(pfb, capi, oracle):
    api = __import__("pfilter_noetic_b200.api", fromlist=["api"])
    m = api.LaserMappingClass(max_map_points=1 << 20, max_points=65536)
    m.init(0.4)
    assert len(m.getMap()) == 0
    m.updateCurrentPointsToMap(np.zeros((0, 4), np.float32), np.eye(4))
    assert m.status == 0 and len(m.getMap()) == 0
    pts = np.array([[1, 2, 3, 9], [1.1, 2.1, 3.1, 9], [40, -60, 2, 9], [500, 0, 0, 9]], np.float32)
    m.updateCurrentPointsToMap(pts, np.eye(4))
    om = oracle.Mapping(0.4)
    om.update(pts, np.eye(4)[:3].reshape(12))
    assert _canon(m.getMap()).tobytes() == _canon(om.get_map()).tobytes()
    assert m.mapping.stats()["dropped"] == 1 == om.dropped()      # 500 m is outside the 5x5x5 block of 50 m cells


def test_mapping_capacity_error(capi):
    m = capi.Mapping(0.4, max_map_points=1000, max_points=65536)
    rng = np.random.default_rng(0)
    pts = (rng.random((50000, 4), dtype=np.float32) - 0.5) * 80
    m.update(pts, np.eye(4)[:3].reshape(12))
    with pytest.raises(capi.PfError) as e:
        m.size()
    assert e.value.status == -3
