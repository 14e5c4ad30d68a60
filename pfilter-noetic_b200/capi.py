"""ctypes binding of the C ABI in include/pfilter_b200.h (libpfilter_b200.so, CUDA sm_100a).

There is no fallback: a missing library or a failing call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PFILTER_B200_LIB selects another build of the same library (kernel variants under build_variants/ while tuning)
LIB_PATH = os.environ.get("PFILTER_B200_LIB") or os.path.join(_HERE, "libpfilter_b200.so")

PF_OK = 0


class PfError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"pfilter_b200 status {status}: {text}")
        self.status = status


class LidarParams(C.Structure):
    _fields_ = [("num_lines", C.c_int32), ("min_distance", C.c_double), ("max_distance", C.c_double),
                ("scan_period", C.c_double)]


class ExtractConfig(C.Structure):
    _fields_ = [("max_points", C.c_int32), ("max_batch", C.c_int32), ("max_ring_points", C.c_int32), ("surf_order", C.c_int32)]


class OdomParams(C.Structure):
    _fields_ = [("map_resolution", C.c_double), ("k_new", C.c_int32), ("theta_p", C.c_float), ("theta_max", C.c_int32),
                ("weight_type", C.c_double), ("max_map_points", C.c_int32), ("max_features", C.c_int32)]


class OdomStats(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("n_edge_ds", "n_surf_ds", "n_edge_res", "n_surf_res", "map_edge", "map_surf",
                                         "passes", "lm_iterations")]


POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("a", "u1")])
assert POINT_DTYPE.itemsize == 16

_lib = None


def lib():
    """Loads libpfilter_b200.so; raises if it has not been built (no CPU fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        _lib.pf_last_error.restype = C.c_char_p
        for name in ("pf_extract_stream", "pf_odom_stream", "pf_mapping_stream"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = C.c_void_p
    return _lib


def check(status):
    if status != PF_OK:
        raise PfError(status, lib().pf_last_error().decode())


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def as_points(a):
    """(n,4) float32 x,y,z,w  ->  contiguous array (bits of w are ignored by the odometry entry points)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def make_points(xyz, r=0, g=0, b=0, a=255):
    out = np.zeros(len(xyz), POINT_DTYPE)
    out["x"], out["y"], out["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    out["r"], out["g"], out["b"], out["a"] = r, g, b, a
    return out


def host_alloc(nbytes):
    p = C.c_void_p()
    check(lib().pf_host_alloc(C.byref(p), C.c_uint64(nbytes)))
    return p


def host_free(p):
    check(lib().pf_host_free(p))


def pinned_array(shape, dtype):
    """numpy array backed by pinned host memory (kept alive by the returned array's base object)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = host_alloc(max(n, 16))
    buf = (C.c_char * max(n, 16)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    return arr, p


class Pc2Layout(C.Structure):
    _fields_ = [("point_step", C.c_uint32), ("row_step", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("off_x", C.c_int32), ("off_y", C.c_int32), ("off_z", C.c_int32), ("off_intensity", C.c_int32),
                ("type_x", C.c_uint8), ("type_y", C.c_uint8), ("type_z", C.c_uint8), ("type_intensity", C.c_uint8), ("is_bigendian", C.c_uint8)]


def pack_pointcloud2(data, point_step, fields, width, height=1, row_step=0, out=None):
    """pf_pack_pointcloud2: PointCloud2 payload -> float32 [n, 4].  fields: {name: (offset, PointField datatype)}; `out` may be a pinned
    [cap, 4] float32 array (pinned_array) -- the scan is packed straight into the buffer the H2D copy reads."""
    buf = np.frombuffer(data, dtype=np.uint8)
    n = int(width) * int(height)
    res = out if out is not None else np.empty((max(n, 1), 4), np.float32)
    assert res.dtype == np.float32 and res.ndim == 2 and res.shape[1] == 4 and res.flags["C_CONTIGUOUS"]
    get = lambda k: fields.get(k, (-1, 7))
    L = Pc2Layout(point_step, row_step, width, height, get("x")[0], get("y")[0], get("z")[0], get("intensity")[0],
                  get("x")[1], get("y")[1], get("z")[1], get("intensity")[1], 0)
    m = C.c_int()
    check(lib().pf_pack_pointcloud2(_vp(buf), C.c_uint64(buf.size), C.byref(L), _vp(res), len(res), C.byref(m)))
    return res[:m.value]


def read_kitti_bin(path, out=None, cap_points=262144):
    """pf_read_kitti_bin: velodyne .bin -> float32 [n, 4] (into `out`, e.g. a pinned array, when given)."""
    res = out if out is not None else np.empty((cap_points, 4), np.float32)
    m = C.c_int()
    check(lib().pf_read_kitti_bin(os.fsencode(str(path)), _vp(res), len(res), C.byref(m)))
    return res[:m.value]


def write_kitti_bin(path, xyzi):
    a = as_points(xyzi)
    check(lib().pf_write_kitti_bin(os.fsencode(str(path)), _vp(a), len(a)))


class Extractor:
    """Handle of pf_extract_* (replaces LaserProcessingClass)."""

    def __init__(self, num_lines=64, min_distance=3.0, max_distance=90.0, max_points=131072, max_batch=1,
                 max_ring_points=0, device=0, surf_order=0):
        self.lidar = LidarParams(num_lines, min_distance, max_distance, 0.1)
        self.cfg = ExtractConfig(max_points, max_batch, max_ring_points, surf_order)
        self.h = C.c_void_p()
        check(lib().pf_extract_create(C.byref(self.lidar), C.byref(self.cfg), device, C.byref(self.h)))
        self.num_lines = num_lines
        self.stride = (max_points + 255) // 256 * 256
        self.edge_stride = 120 * num_lines
        self.max_batch = max_batch

    def close(self):
        if self.h:
            lib().pf_extract_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, xyzi, want_label=True):
        a = as_points(xyzi)
        n = a.shape[0]
        edge = np.empty((self.edge_stride, 4), np.float32)
        surf = np.empty((max(n, 1), 4), np.float32)
        label = np.zeros(max(n, 1), np.uint8) if want_label else None
        ne, ns = C.c_int(), C.c_int()
        check(lib().pf_extract_run(self.h, _vp(a), n, _vp(edge), C.byref(ne), _vp(surf), C.byref(ns), _vp(label)))
        return edge[:ne.value].copy(), surf[:ns.value].copy(), (label[:n] if want_label else None)

    def run_batch(self, scans, want_label=True):
        b = len(scans)
        stride = self.stride
        x = np.zeros((b, stride, 4), np.float32)
        n = np.zeros(b, np.int32)
        for i, s in enumerate(scans):
            n[i] = len(s)
            x[i, :len(s)] = s
        edge = np.empty((b, self.edge_stride, 4), np.float32)
        surf = np.empty((b, stride, 4), np.float32)
        label = np.zeros((b, stride), np.uint8) if want_label else None
        ne = np.zeros(b, np.int32)
        ns = np.zeros(b, np.int32)
        check(lib().pf_extract_run_batch(self.h, _vp(x), _vp(n), b, stride, _vp(edge), _vp(ne), self.edge_stride, _vp(surf),
                                         _vp(ns), _vp(label)))
        return [(edge[i, :ne[i]].copy(), surf[i, :ns[i]].copy(), None if label is None else label[i, :n[i]].copy())
                for i in range(b)]

    def run_batch_device(self, d_xyzi, d_n, batch, stride, d_edge, d_n_edge, d_surf, d_n_surf, d_label=0):
        """All arguments are raw device pointers (ints); enqueues on the handle's stream without synchronising."""
        check(lib().pf_extract_run_batch_device(self.h, C.c_void_p(d_xyzi), C.c_void_p(d_n), batch, stride, C.c_void_p(d_edge),
                                                C.c_void_p(d_n_edge), self.edge_stride, C.c_void_p(d_surf),
                                                C.c_void_p(d_n_surf), C.c_void_p(d_label)))

    def sync(self):
        check(lib().pf_extract_sync(self.h))

    @property
    def stream(self):
        return lib().pf_extract_stream(self.h)

    @property
    def launches(self):
        v = C.c_uint64()
        check(lib().pf_extract_kernel_launches(self.h, C.byref(v)))
        return v.value


# ------------------------------------------------------------------------------------------------------------
# stage taps
# ------------------------------------------------------------------------------------------------------------
def _pts(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == POINT_DTYPE, a.dtype
    return a


def voxel_downsample(pts, leaf, device=0):
    p = _pts(pts)
    out = np.empty(max(len(p), 1), POINT_DTYPE)
    n = C.c_int()
    check(lib().pf_voxel_downsample(device, _vp(p), len(p), C.c_float(leaf), _vp(out), C.byref(n)))
    return out[:n.value].copy()


def map_update(pts, center, leaf, k_new, theta_p, theta_max, device=0):
    p = _pts(pts)
    out = np.empty(max(len(p), 1), POINT_DTYPE)
    n = C.c_int()
    c = np.ascontiguousarray(center, np.float64)
    check(lib().pf_map_update(device, _vp(p), len(p), _vp(c), C.c_float(leaf), k_new, C.c_float(theta_p), theta_max, _vp(out),
                              C.byref(n)))
    return out[:n.value].copy()


def map_merge(sorted_map, extra, center, leaf, k_new, theta_p, theta_max, device=0):
    """pf_map_merge: returns (map, n_sorted)."""
    a, b = _pts(sorted_map), _pts(extra)
    out = np.empty(max(len(a) + len(b), 1), POINT_DTYPE)
    n, ns = C.c_int(), C.c_int()
    c = np.asarray(center, np.float64)
    check(lib().pf_map_merge(device, _vp(a), len(a), _vp(b), len(b), _vp(c), C.c_float(leaf), k_new, C.c_float(theta_p), theta_max,
                             _vp(out), len(out), C.byref(n), C.byref(ns)))
    return out[:n.value].copy(), ns.value


def map_merge_timed(sorted_map, extra, center, leaf, k_new, theta_p, theta_max, reps=5, device=0):
    """pf_map_merge_timed: returns dict(n_out, n_sorted, ms_total, ms_stream)."""
    a, b = _pts(sorted_map), _pts(extra)
    n, ns, mt, ms = C.c_int(), C.c_int(), C.c_float(), C.c_float()
    c = np.asarray(center, np.float64)
    check(lib().pf_map_merge_timed(device, _vp(a), len(a), _vp(b), len(b), _vp(c), C.c_float(leaf), k_new, C.c_float(theta_p), theta_max,
                                   reps, C.byref(n), C.byref(ns), C.byref(mt), C.byref(ms)))
    return {"n_out": n.value, "n_sorted": ns.value, "ms_total": mt.value, "ms_stream": ms.value}


def knn5_timed(map_pts, queries_xyz4, reps=5, device=0):
    m = _pts(map_pts)
    q = np.ascontiguousarray(queries_xyz4, np.float32)
    idx = np.empty((max(len(q), 1), 5), np.int32)
    d2 = np.empty((max(len(q), 1), 5), np.float32)
    mb, mq = C.c_float(), C.c_float()
    check(lib().pf_knn5_timed(device, _vp(m), len(m), _vp(q), len(q), _vp(idx), _vp(d2), reps, C.byref(mb), C.byref(mq)))
    return idx[:len(q)], d2[:len(q)], mb.value, mq.value


def knn5(map_pts, queries_xyz4, device=0):
    m = _pts(map_pts)
    q = np.ascontiguousarray(queries_xyz4, np.float32)
    assert q.ndim == 2 and q.shape[1] == 4
    idx = np.empty((max(len(q), 1), 5), np.int32)
    d2 = np.empty((max(len(q), 1), 5), np.float32)
    check(lib().pf_knn5(device, _vp(m), len(m), _vp(q), len(q), _vp(idx), _vp(d2)))
    return idx[:len(q)], d2[:len(q)]


def associate(kind, map_pts, queries, pose, k_new, theta_p, theta_max, device=0):
    """One association pass; returns (map_after, queries_after, flag, geom8)."""
    m = _pts(map_pts).copy()
    q = _pts(queries).copy()
    flag = np.zeros(max(len(q), 1), np.uint8)
    geom = np.zeros((max(len(q), 1), 8), np.float64)
    pose = np.ascontiguousarray(pose, np.float64)
    check(lib().pf_associate(device, kind, _vp(m), len(m), _vp(q), len(q), _vp(pose), k_new, C.c_float(theta_p), theta_max, _vp(flag),
                             _vp(geom)))
    return m, q, flag[:len(q)], geom[:len(q)]


def eval_normal_eq(pose, edge9, surf7, device=0):
    pose = np.ascontiguousarray(pose, np.float64)
    e = np.ascontiguousarray(edge9, np.float64).reshape(-1, 9)
    s = np.ascontiguousarray(surf7, np.float64).reshape(-1, 7)
    H = np.zeros(21); g = np.zeros(6); cost = C.c_double()
    check(lib().pf_eval_normal_eq(device, _vp(pose), _vp(e), len(e), _vp(s), len(s), _vp(H), _vp(g), C.byref(cost)))
    return H, g, cost.value


def eval_normal_eq_timed(pose, edge9, surf7, reps=5, device=0):
    """pf_eval_normal_eq_timed (grid-wide streaming kernel): returns (H21, g6, cost, ms_kernel)."""
    pose = np.ascontiguousarray(pose, np.float64)
    e = np.ascontiguousarray(edge9, np.float64).reshape(-1, 9)
    s = np.ascontiguousarray(surf7, np.float64).reshape(-1, 7)
    H = np.zeros(21); g = np.zeros(6); cost = C.c_double(); ms = C.c_float()
    check(lib().pf_eval_normal_eq_timed(device, _vp(pose), _vp(e), len(e), _vp(s), len(s), reps, _vp(H), _vp(g), C.byref(cost), C.byref(ms)))
    return H, g, cost.value, ms.value


def lm_solve(pose, edge9, surf7, device=0):
    x = np.array(pose, np.float64)
    e = np.ascontiguousarray(edge9, np.float64).reshape(-1, 9)
    s = np.ascontiguousarray(surf7, np.float64).reshape(-1, 7)
    it = C.c_int(); cost = C.c_double()
    check(lib().pf_lm_solve(device, _vp(x), _vp(e), len(e), _vp(s), len(s), C.byref(it), C.byref(cost)))
    return x, it.value, cost.value


class Odometry:
    """Handle of pf_odom_* (replaces Odom_ES_EstimationClass)."""

    def __init__(self, map_resolution=0.4, k_new=0, theta_p=0.4, theta_max=75, weight_type=0.0, max_map_points=0, max_features=0,
                 device=0):
        self.prm = OdomParams(map_resolution, k_new, theta_p, theta_max, weight_type, max_map_points, max_features)
        self.h = C.c_void_p()
        check(lib().pf_odom_create(C.byref(self.prm), device, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().pf_odom_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def init_map(self, edge4, surf4):
        e, s = as_points(edge4), as_points(surf4)
        check(lib().pf_odom_init_map(self.h, _vp(e), len(e), _vp(s), len(s)))

    def update(self, edge4, surf4):
        e, s = as_points(edge4), as_points(surf4)
        pose = np.zeros(7)
        check(lib().pf_odom_update(self.h, _vp(e), len(e), _vp(s), len(s), _vp(pose)))
        return pose

    def process_extracted(self, extractor):
        pose = np.zeros(7)
        check(lib().pf_odom_process_extracted(self.h, extractor.h, _vp(pose)))
        return pose

    def pose(self):
        pose = np.zeros(7)
        check(lib().pf_odom_get_pose(self.h, _vp(pose)))
        return pose

    def map_part(self, which):
        n = C.c_int()
        check(lib().pf_odom_map_size(self.h, which, C.byref(n)))
        out = np.empty(max(n.value, 1), POINT_DTYPE)
        check(lib().pf_odom_get_map_part(self.h, which, _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def get_map(self):
        ne, ns = C.c_int(), C.c_int()
        check(lib().pf_odom_map_size(self.h, 0, C.byref(ne)))
        check(lib().pf_odom_map_size(self.h, 1, C.byref(ns)))
        out = np.empty(max(ne.value + ns.value, 1), POINT_DTYPE)
        n = C.c_int()
        check(lib().pf_odom_get_map(self.h, _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()

    def iter_poses(self):
        out = np.zeros((16, 7))
        n = C.c_int()
        check(lib().pf_odom_get_iter_poses(self.h, _vp(out), 16, C.byref(n)))
        return out[:n.value].copy()

    def stats(self):
        s = OdomStats()
        check(lib().pf_odom_get_stats(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in OdomStats._fields_}

    @property
    def stream(self):
        return lib().pf_odom_stream(self.h)

    @property
    def launches(self):
        v = C.c_uint64()
        check(lib().pf_odom_kernel_launches(self.h, C.byref(v)))
        return v.value

    @property
    def graph_captures(self):
        v = C.c_int()
        check(lib().pf_odom_graph_captures(self.h, C.byref(v)))
        return v.value


class OdometryBPF(Odometry):
    """Handle of pf_odom_bpf_* (replaces Odom_BPF_EstimationClass): beam / pillar / facade feature kinds."""

    def __init__(self, map_resolution=0.4, k_new=0, theta_p=0.4, theta_max=75, weight_type=0.0, max_map_points=0, max_features=0,
                 device=0):
        self.prm = OdomParams(map_resolution, k_new, theta_p, theta_max, weight_type, max_map_points, max_features)
        self.h = C.c_void_p()
        check(lib().pf_odom_bpf_create(C.byref(self.prm), device, C.byref(self.h)))

    def init_map(self, beam4, pillar4, facade4):
        b, p, f = as_points(beam4), as_points(pillar4), as_points(facade4)
        check(lib().pf_odom_bpf_init_map(self.h, _vp(b), len(b), _vp(p), len(p), _vp(f), len(f)))

    def update(self, beam4, pillar4, facade4):
        b, p, f = as_points(beam4), as_points(pillar4), as_points(facade4)
        pose = np.zeros(7)
        check(lib().pf_odom_bpf_update(self.h, _vp(b), len(b), _vp(p), len(p), _vp(f), len(f), _vp(pose)))
        return pose

    def get_map(self):
        sizes = []
        for which in range(3):
            n = C.c_int()
            check(lib().pf_odom_map_size(self.h, which, C.byref(n)))
            sizes.append(n.value)
        out = np.empty(max(sum(sizes), 1), POINT_DTYPE)
        n = C.c_int()
        check(lib().pf_odom_get_map(self.h, _vp(out), len(out), C.byref(n)))
        return out[:n.value].copy()


def pose_to_rt(pose7):
    """[qx qy qz qw tx ty tz] -> row-major 3x4 [R | t] (Eigen::Quaterniond::toRotationMatrix, double)."""
    x, y, z, w = (float(v) for v in pose7[:4])
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz, txx, txy, txz, tyy, tyz, tzz = tx * w, ty * w, tz * w, tx * x, ty * x, tz * x, ty * y, tz * y, tz * z
    return np.array([1 - (tyy + tzz), txy - twz, txz + twy, pose7[4],
                     txy + twz, 1 - (txx + tzz), tyz - twx, pose7[5],
                     txz - twy, tyz + twx, 1 - (txx + tyy), pose7[6]], np.float64)


class Mapping:
    """Handle of pf_mapping_* (replaces LaserMappingClass)."""

    def __init__(self, map_resolution=0.4, max_map_points=0, max_points=0, device=0):
        self.h = C.c_void_p()
        check(lib().pf_mapping_create(C.c_double(map_resolution), max_map_points, max_points, device, C.byref(self.h)))

    def close(self):
        if self.h:
            lib().pf_mapping_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def update(self, xyzi, rt12):
        a = as_points(xyzi)
        rt = np.ascontiguousarray(rt12, np.float64).reshape(12)
        check(lib().pf_mapping_update(self.h, _vp(a), len(a), _vp(rt)))

    def size(self):
        n = C.c_int()
        check(lib().pf_mapping_map_size(self.h, C.byref(n)))
        return n.value

    def get_map(self):
        n = self.size()
        out = np.empty((max(n, 1), 4), np.float32)
        m = C.c_int()
        check(lib().pf_mapping_get_map(self.h, _vp(out), len(out), C.byref(m)))
        return out[:m.value].copy()

    def stats(self):
        ns, dr, la = C.c_int(), C.c_longlong(), C.c_uint64()
        check(lib().pf_mapping_stats(self.h, C.byref(ns), C.byref(dr), C.byref(la)))
        return {"n_sorted": ns.value, "dropped": dr.value, "launches": la.value}


def frame_submit(extractor, odometry, xyzi):
    """pf_frame_submit: asynchronous; returns the frame id.  `xyzi` must stay alive (pinned) until frame_wait(frame id)."""
    a = as_points(xyzi)
    fid = C.c_longlong()
    check(lib().pf_frame_submit(extractor.h, odometry.h, _vp(a), len(a), C.byref(fid)))
    return fid.value


def frame_wait(odometry, frame_id):
    pose = np.zeros(7)
    check(lib().pf_frame_wait(odometry.h, C.c_longlong(frame_id), _vp(pose)))
    return pose


def frame_process(extractor, odometry, xyzi):
    """pf_frame_process: H2D scan -> extract -> (init | update) -> pose."""
    a = as_points(xyzi)
    pose = np.zeros(7)
    check(lib().pf_frame_process(extractor.h, odometry.h, _vp(a), len(a), _vp(pose)))
    return pose
