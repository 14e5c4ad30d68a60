"""CPU: scan packing helpers (KITTI .bin, PointCloud2 payloads) produce the ABI's float4 layout."""
import numpy as np


def test_kitti_roundtrip(pfb, tmp_path):
    io = __import__("pfilter_noetic_b200.io", fromlist=["io"])
    rng = np.random.default_rng(0)
    a = rng.standard_normal((1000, 4)).astype(np.float32)
    io.write_kitti_bin(tmp_path / "000000.bin", a)
    b = io.read_kitti_bin(tmp_path / "000000.bin")
    assert b.dtype == np.float32 and b.flags["C_CONTIGUOUS"] and np.array_equal(a, b)


def test_pointcloud2_roundtrip_and_missing_intensity(pfb):
    io = __import__("pfilter_noetic_b200.io", fromlist=["io"])
    rng = np.random.default_rng(1)
    a = rng.standard_normal((257, 4)).astype(np.float32)
    data, step, fields = io.xyzi_to_pointcloud2(a)
    assert step == 32 and len(data) == 32 * 257
    assert np.array_equal(io.pointcloud2_to_xyzi(data, step, fields, 257), a)
    del fields["intensity"]
    b = io.pointcloud2_to_xyzi(data, step, fields, 257)
    assert np.array_equal(b[:, :3], a[:, :3]) and not b[:, 3].any()
    # velodyne driver layout: x y z f32, intensity f32 at 12, ring u16 at 16, point_step 22 rounded to 32
    rec = np.zeros(10, dtype=np.dtype({"names": ["x", "y", "z", "intensity", "ring"], "formats": ["<f4"] * 4 + ["<u2"], "offsets": [0, 4, 8, 12, 16], "itemsize": 32}))
    rec["x"], rec["y"], rec["z"], rec["intensity"], rec["ring"] = a[:10, 0], a[:10, 1], a[:10, 2], a[:10, 3], np.arange(10)
    c = io.pointcloud2_to_xyzi(rec.tobytes(), 32, {"x": (0, 7), "y": (4, 7), "z": (8, 7), "intensity": (12, 7), "ring": (16, 4)}, 10)
    assert np.array_equal(c, a[:10])
