"""Wire / disk side of the hot path: the 16-byte float4 scan layout of the C ABI from the formats
the reference's callers hold.

* KITTI odometry ``velodyne/*.bin``: little-endian float32 x, y, z, reflectance -- already the ABI layout (the reference reaches
  it through a bag / PointCloud2 publisher, launch/pfilter_kitti.launch).
* ``sensor_msgs/PointCloud2`` as the reference's nodes receive it (src/laserProcessingNode.cpp:52-63 -> pcl::fromROSMsg into
  PointXYZI): raw byte buffer + point_step + field offsets, no ROS import needed.
"""
import numpy as np

_PC2_DTYPES = {1: "i1", 2: "u1", 3: "<i2", 4: "<u2", 5: "<i4", 6: "<u4", 7: "<f4", 8: "<f8"}   # sensor_msgs/PointField datatype codes


def read_kitti_bin(path):
    """KITTI velodyne scan -> float32 [n, 4] (x, y, z, intensity), ready for pf_extract_run / pf_frame_process."""
    a = np.fromfile(path, dtype="<f4")
    if a.size % 4:
        raise ValueError(f"{path}: {a.size} floats is not a multiple of 4")
    return np.ascontiguousarray(a.reshape(-1, 4))


def write_kitti_bin(path, xyzi):
    np.ascontiguousarray(xyzi, dtype="<f4").reshape(-1, 4).tofile(path)


def pointcloud2_to_xyzi(data, point_step, fields, width, height=1, row_step=None, is_bigendian=False, out=None):
    """PointCloud2 payload -> float32 [n, 4].

    ``fields``: {name: (offset, datatype)} with the PointField datatype codes; x, y, z are required, ``intensity`` is optional
    (0 when absent, like pcl::fromROSMsg leaves it).  ``out`` may be a pinned [n, 4] float32 array (capi.pinned_array) so
    the scan is packed straight into the buffer the H2D copy reads."""
    if is_bigendian:
        raise ValueError("big-endian PointCloud2 payloads are not supported")
    n = int(width) * int(height)
    row_step = int(row_step) if row_step else int(point_step) * int(width)
    buf = np.frombuffer(data, dtype=np.uint8)
    if height > 1 and row_step != point_step * width:      # padded rows
        buf = buf[:row_step * height].reshape(height, row_step)[:, :point_step * width].reshape(-1)
    pts = buf[:n * point_step].reshape(n, point_step)
    res = out if out is not None else np.empty((n, 4), np.float32)
    if res.shape != (n, 4) or res.dtype != np.float32:
        raise ValueError("out must be float32 [n, 4]")
    for col, name in enumerate(("x", "y", "z", "intensity")):
        if name not in fields:
            if name == "intensity":
                res[:, 3] = 0
                continue
            raise KeyError(f"PointCloud2 has no field {name!r}")
        off, dt = fields[name]
        dt = np.dtype(_PC2_DTYPES[int(dt)])
        res[:, col] = np.ascontiguousarray(pts[:, off:off + dt.itemsize]).view(dt).reshape(n)
    return res


def xyzi_to_pointcloud2(xyzi):
    """float32 [n, 4] -> (data bytes, point_step, fields) in the layout pcl::toROSMsg gives PointXYZI clouds
    (x, y, z at 0/4/8, intensity at 16, point_step 32: PCL pads PointXYZI to 32 bytes)."""
    a = np.ascontiguousarray(xyzi, np.float32).reshape(-1, 4)
    rec = np.zeros(len(a), dtype=np.dtype({"names": ["x", "y", "z", "intensity"], "formats": ["<f4"] * 4, "offsets": [0, 4, 8, 16], "itemsize": 32}))
    rec["x"], rec["y"], rec["z"], rec["intensity"] = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
    return rec.tobytes(), 32, {"x": (0, 7), "y": (4, 7), "z": (8, 7), "intensity": (16, 7)}
