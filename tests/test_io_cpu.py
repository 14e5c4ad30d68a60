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


def test_c_abi_scan_packing_matches_python(pfb, tmp_path):
    """pf_pack_pointcloud2 / pf_read_kitti_bin / pf_write_kitti_bin (host-only entry points of the C ABI) against the numpy helpers."""
    import pytest
    io = __import__("pfilter_noetic_b200.io", fromlist=["io"])
    capi = pfb.capi
    rng = np.random.default_rng(2)
    a = rng.standard_normal((1031, 4)).astype(np.float32)
    capi.write_kitti_bin(tmp_path / "a.bin", a)
    assert np.array_equal(io.read_kitti_bin(tmp_path / "a.bin"), a)
    io.write_kitti_bin(tmp_path / "b.bin", a)
    assert np.array_equal(capi.read_kitti_bin(tmp_path / "b.bin"), a)
    with pytest.raises(capi.PfError) as e:
        capi.read_kitti_bin(tmp_path / "b.bin", cap_points=1000)
    assert e.value.status == -3
    with open(tmp_path / "c.bin", "wb") as f:
        f.write(a.tobytes()[:-3])
    with pytest.raises(capi.PfError):
        capi.read_kitti_bin(tmp_path / "c.bin")
    with pytest.raises(capi.PfError):
        capi.read_kitti_bin(tmp_path / "missing.bin")
    # PointCloud2: the pcl::toROSMsg layout, a velodyne-driver layout with mixed types, padded rows, missing intensity
    data, step, fields = io.xyzi_to_pointcloud2(a)
    assert np.array_equal(capi.pack_pointcloud2(data, step, fields, len(a)), a)
    f2 = dict(fields); del f2["intensity"]
    b = capi.pack_pointcloud2(data, step, f2, len(a))
    assert np.array_equal(b[:, :3], a[:, :3]) and not b[:, 3].any()
    rec = np.zeros(12, dtype=np.dtype({"names": ["x", "y", "z", "intensity", "ring"], "formats": ["<f8", "<f4", "<f4", "<u2", "<u2"], "offsets": [0, 8, 12, 16, 18], "itemsize": 24}))
    rec["x"], rec["y"], rec["z"], rec["intensity"] = a[:12, 0], a[:12, 1], a[:12, 2], np.arange(12) * 7
    fl = {"x": (0, 8), "y": (8, 7), "z": (12, 7), "intensity": (16, 4)}
    c = capi.pack_pointcloud2(rec.tobytes(), 24, fl, 12)
    assert np.array_equal(c, io.pointcloud2_to_xyzi(rec.tobytes(), 24, fl, 12))
    rows = np.zeros((3, 4 * 32 + 16), np.uint8)                     # height 3, width 4, row padding of 16 bytes
    rows[:, :128] = np.frombuffer(io.xyzi_to_pointcloud2(a[:12])[0], np.uint8).reshape(3, 128)
    d = capi.pack_pointcloud2(rows.tobytes(), 32, fields, 4, height=3, row_step=144)
    assert np.array_equal(d, a[:12])
    with pytest.raises(capi.PfError):                                # payload shorter than the layout says
        capi.pack_pointcloud2(data[:100], step, fields, len(a))
    with pytest.raises(capi.PfError):
        capi.pack_pointcloud2(data, step, {"y": (4, 7), "z": (8, 7)}, len(a))
