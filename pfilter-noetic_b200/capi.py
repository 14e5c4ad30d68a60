"""ctypes binding of the C ABI in include/pfilter_b200.h (libpfilter_b200.so, CUDA sm_100a).

There is no fallback: a missing library or a failing call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpfilter_b200.so")

PF_OK = 0


class PfError(RuntimeError):
    def __init__(self, status, text):
        super().__init__(f"pfilter_b200 status {status}: {text}")
        self.status = status


class LidarParams(C.Structure):
    _fields_ = [("num_lines", C.c_int32), ("min_distance", C.c_double), ("max_distance", C.c_double),
                ("scan_period", C.c_double)]


class ExtractConfig(C.Structure):
    _fields_ = [("max_points", C.c_int32), ("max_batch", C.c_int32), ("max_ring_points", C.c_int32)]


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
        for name in ("pf_extract_stream", "pf_odom_stream"):
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


class Extractor:
    """Handle of pf_extract_* (replaces LaserProcessingClass)."""

    def __init__(self, num_lines=64, min_distance=3.0, max_distance=90.0, max_points=131072, max_batch=1,
                 max_ring_points=0, device=0):
        self.lidar = LidarParams(num_lines, min_distance, max_distance, 0.1)
        self.cfg = ExtractConfig(max_points, max_batch, max_ring_points)
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
