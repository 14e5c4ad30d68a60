"""ORACLE / TEST INFRASTRUCTURE ONLY.  ctypes access to oracle/liboracle.so (CPU restatement) and
oracle/_ref/libpf_ref_extract.so (the real reference extraction sources compiled in place).
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_o = None
_r = None


def lib():
    global _o
    if _o is None:
        _o = C.CDLL(os.path.join(_HERE, "liboracle.so"))
    return _o


def ref_lib():
    global _r
    if _r is None:
        _r = C.CDLL(os.path.join(_HERE, "_ref", "libpf_ref_extract.so"))
    return _r


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libpf_ref_extract.so"))


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    assert a.ndim == 2 and a.shape[1] == 4
    return a


def ref_extract(xyzi, num_lines=64, min_d=3.0, max_d=90.0):
    """Real reference: returns (edge_idx, surf_idx) input indices in the reference's emission order."""
    a = _f32(xyzi)
    n = a.shape[0]
    e = np.empty(n, np.int32); s = np.empty(n, np.int32)
    ne = C.c_int(); ns = C.c_int()
    rc = ref_lib().pfref_extract(a.ctypes.data_as(C.c_void_p), n, num_lines, C.c_double(min_d), C.c_double(max_d),
                                 e.ctypes.data_as(C.c_void_p), C.byref(ne), s.ctypes.data_as(C.c_void_p), C.byref(ns))
    assert rc == 0
    return e[:ne.value].copy(), s[:ns.value].copy()


def ref_extract_time(xyzi, reps, num_lines=64, min_d=3.0, max_d=90.0):
    import time
    a = _f32(xyzi)
    ne = C.c_int(); ns = C.c_int()
    t0 = time.perf_counter()
    ref_lib().pfref_extract_time(a.ctypes.data_as(C.c_void_p), a.shape[0], num_lines, C.c_double(min_d), C.c_double(max_d),
                                 reps, C.byref(ne), C.byref(ns))
    return (time.perf_counter() - t0) / reps


def extract(xyzi, num_lines=64, min_d=3.0, max_d=90.0, order=1):
    """Restatement: returns dict(edge_idx, surf_idx, label, ring)."""
    a = _f32(xyzi)
    n = a.shape[0]
    e = np.empty(max(n, 1), np.int32); s = np.empty(max(n, 1), np.int32)
    label = np.zeros(max(n, 1), np.uint8); ring = np.empty(max(n, 1), np.int32)
    ne = C.c_int(); ns = C.c_int()
    rc = lib().pforacle_extract(a.ctypes.data_as(C.c_void_p), n, num_lines, C.c_double(min_d), C.c_double(max_d), order,
                                e.ctypes.data_as(C.c_void_p), C.byref(ne), s.ctypes.data_as(C.c_void_p), C.byref(ns),
                                label.ctypes.data_as(C.c_void_p), ring.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return dict(edge_idx=e[:ne.value].copy(), surf_idx=s[:ns.value].copy(), label=label[:n].copy(), ring=ring[:n].copy())


# ------------------------------------------------------------------------------------------------------------
# odometry / map restatement (oracle_odom.cpp)
# ------------------------------------------------------------------------------------------------------------
POINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("g", "u1"), ("b", "u1"), ("a", "u1")])


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _pts(a):
    a = np.ascontiguousarray(a)
    assert a.dtype == POINT_DTYPE
    return a


def voxel_downsample(pts, leaf):
    p = _pts(pts)
    out = np.empty(max(len(p), 1), POINT_DTYPE)
    n = C.c_int()
    lib().pforacle_voxel_downsample(_vp(p), len(p), C.c_float(leaf), _vp(out), C.byref(n))
    return out[:n.value].copy()


def map_update(pts, center, leaf, k_new, theta_p, theta_max):
    p = _pts(pts)
    out = np.empty(max(len(p), 1), POINT_DTYPE)
    n = C.c_int()
    c = np.asarray(center, np.float64)
    lib().pforacle_map_update(_vp(p), len(p), _vp(c), C.c_float(leaf), k_new, C.c_float(theta_p), theta_max, _vp(out), C.byref(n))
    return out[:n.value].copy()


def knn5(map_pts, queries_xyz4, mode=1):
    m = _pts(map_pts)
    q = np.ascontiguousarray(queries_xyz4, np.float32)
    idx = np.empty((len(q), 5), np.int32)
    d2 = np.empty((len(q), 5), np.float32)
    lib().pforacle_knn5(_vp(m), len(m), _vp(q), len(q), mode, _vp(idx), _vp(d2))
    return idx, d2


def knn5_timed(map_pts, queries_xyz4):
    """kd-tree (FLANN KDTreeSingleIndex-like, leaf 15) build and query timed separately: (idx, d2, s_build, s_query)."""
    m = _pts(map_pts)
    q = np.ascontiguousarray(queries_xyz4, np.float32)
    idx = np.empty((len(q), 5), np.int32)
    d2 = np.empty((len(q), 5), np.float32)
    tb, tq = C.c_double(), C.c_double()
    lib().pforacle_knn5_timed(_vp(m), len(m), _vp(q), len(q), _vp(idx), _vp(d2), C.byref(tb), C.byref(tq))
    return idx, d2, tb.value, tq.value


def associate(kind, map_pts, queries, pose, k_new, theta_p, theta_max):
    """Returns (map_after, queries_after, flag, geom8)."""
    m = _pts(map_pts).copy()
    q = _pts(queries).copy()
    flag = np.zeros(len(q), np.uint8)
    geom = np.zeros((len(q), 8), np.float64)
    pose = np.ascontiguousarray(pose, np.float64)
    lib().pforacle_associate(kind, _vp(m), len(m), _vp(q), len(q), _vp(pose), k_new, C.c_float(theta_p), theta_max, _vp(flag), _vp(geom))
    return m, q, flag, geom


def eval_normal_eq(pose, edge9, surf7):
    pose = np.ascontiguousarray(pose, np.float64)
    e = np.ascontiguousarray(edge9, np.float64).reshape(-1, 9)
    s = np.ascontiguousarray(surf7, np.float64).reshape(-1, 7)
    H = np.zeros(21); g = np.zeros(6); cost = C.c_double()
    lib().pforacle_eval_normal_eq(_vp(pose), _vp(e), len(e), _vp(s), len(s), _vp(H), _vp(g), C.byref(cost))
    return H, g, cost.value


def lm_solve(pose, edge9, surf7):
    x = np.array(pose, np.float64)
    e = np.ascontiguousarray(edge9, np.float64).reshape(-1, 9)
    s = np.ascontiguousarray(surf7, np.float64).reshape(-1, 7)
    it = C.c_int(); cost = C.c_double()
    lib().pforacle_lm_solve(_vp(x), _vp(e), len(e), _vp(s), len(s), C.byref(it), C.byref(cost))
    return x, it.value, cost.value


def lm_solve_ex(pose, edge9, surf7, max_iter=200, ftol=1e-15):
    """The oracle's LM state machine run to convergence (iteration cap and function tolerance opened up)."""
    x = np.array(pose, np.float64)
    e = np.ascontiguousarray(edge9, np.float64).reshape(-1, 9)
    s = np.ascontiguousarray(surf7, np.float64).reshape(-1, 7)
    it = C.c_int(); cost = C.c_double()
    lib().pforacle_lm_solve_ex(_vp(x), _vp(e), len(e), _vp(s), len(s), max_iter, C.c_double(ftol), C.byref(it), C.byref(cost))
    return x, it.value, cost.value


def set_sort_mode(literal):
    """0: stable voxel sort (convention shared with the GPU kernels); 1: the reference's literal (unstable) std::sort.  Returns the old mode."""
    return lib().pforacle_set_sort_mode(int(literal))


def se3_plus(x, d):
    x = np.ascontiguousarray(x, np.float64); d = np.ascontiguousarray(d, np.float64)
    out = np.zeros(7)
    lib().pforacle_se3_plus(_vp(x), _vp(d), _vp(out))
    return out


def eig3(A):
    A = np.ascontiguousarray(A, np.float64)
    w = np.zeros(3); V = np.zeros((3, 3))
    lib().pforacle_eig3(_vp(A), _vp(w), _vp(V))
    return w, V


def qr_solve(A, b):
    A = np.ascontiguousarray(A, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(3)
    lib().pforacle_qr_solve(_vp(A), _vp(b), A.shape[0], _vp(x))
    return x


def residual(kind, geom, pose):
    geom = np.ascontiguousarray(geom, np.float64); pose = np.ascontiguousarray(pose, np.float64)
    r = C.c_double(); J = np.zeros(6)
    lib().pforacle_residual(kind, _vp(geom), _vp(pose), C.byref(r), _vp(J))
    return r.value, J


class Odom:
    """Restated Odom_ES_EstimationClass (CPU)."""

    def __init__(self, map_resolution=0.4, k_new=0, theta_p=0.4, theta_max=75, weight_type=0.0):
        L = lib()
        L.pforacle_odom_create.restype = C.c_void_p
        self.h = C.c_void_p(L.pforacle_odom_create(C.c_double(map_resolution), k_new, C.c_float(theta_p), theta_max, C.c_double(weight_type)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().pforacle_odom_destroy(self.h)
            self.h = None

    def init_map(self, edge4, surf4):
        e = np.ascontiguousarray(edge4, np.float32); s = np.ascontiguousarray(surf4, np.float32)
        lib().pforacle_odom_init_map(self.h, _vp(e), len(e), _vp(s), len(s))

    def update(self, edge4, surf4):
        e = np.ascontiguousarray(edge4, np.float32); s = np.ascontiguousarray(surf4, np.float32)
        pose = np.zeros(7)
        lib().pforacle_odom_update(self.h, _vp(e), len(e), _vp(s), len(s), _vp(pose))
        return pose

    def get_map(self, which):
        n = lib().pforacle_odom_map_size(self.h, which)
        out = np.empty(max(n, 1), POINT_DTYPE)
        lib().pforacle_odom_get_map(self.h, which, _vp(out))
        return out[:n].copy()

    def iter_poses(self):
        out = np.zeros((16, 7))
        n = lib().pforacle_odom_iter_poses(self.h, _vp(out), 16)
        return out[:n].copy()

    def stats(self):
        out = np.zeros(8, np.int32)
        lib().pforacle_odom_stats(self.h, _vp(out))
        return dict(zip(("n_edge_ds", "n_surf_ds", "n_edge_res", "n_surf_res", "map_edge", "map_surf", "passes", "lm_iterations"),
                        out.tolist()))


class Mapping:
    """Restated LaserMappingClass (CPU), oracle_mapping.cpp."""

    def __init__(self, map_resolution=0.4):
        L = lib()
        L.pforacle_mapping_create.restype = C.c_void_p
        L.pforacle_mapping_size.restype = C.c_long
        L.pforacle_mapping_dropped.restype = C.c_long
        self.h = C.c_void_p(L.pforacle_mapping_create(C.c_double(map_resolution)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().pforacle_mapping_destroy(self.h)
            self.h = None

    def update(self, xyzi, rt12):
        a = _f32(xyzi)
        rt = np.ascontiguousarray(rt12, np.float64).reshape(12)
        lib().pforacle_mapping_update(self.h, _vp(a), len(a), _vp(rt))

    def dropped(self):
        return int(lib().pforacle_mapping_dropped(self.h))

    def get_map(self):
        n = int(lib().pforacle_mapping_size(self.h))
        out = np.empty((max(n, 1), 4), np.float32)
        lib().pforacle_mapping_get_map(self.h, _vp(out))
        return out[:n].copy()


class OdomBPF:
    """Restated Odom_BPF_EstimationClass (CPU): beam, pillar (line residuals) and facade (plane residuals)."""

    def __init__(self, map_resolution=0.4, k_new=0, theta_p=0.4, theta_max=75, weight_type=0.0):
        L = lib()
        L.pforacle_bpf_create.restype = C.c_void_p
        self.h = C.c_void_p(L.pforacle_bpf_create(C.c_double(map_resolution), k_new, C.c_float(theta_p), theta_max, C.c_double(weight_type)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().pforacle_bpf_destroy(self.h)
            self.h = None

    def _frame(self, init, beam, pillar, facade):
        a = [np.ascontiguousarray(x, np.float32).reshape(-1, 4) for x in (beam, pillar, facade)]
        pose = np.zeros(7)
        lib().pforacle_bpf_frame(self.h, init, _vp(a[0]), len(a[0]), _vp(a[1]), len(a[1]), _vp(a[2]), len(a[2]), _vp(pose))
        return pose

    def init_map(self, beam, pillar, facade):
        self._frame(1, beam, pillar, facade)

    def update(self, beam, pillar, facade):
        return self._frame(0, beam, pillar, facade)

    def get_map(self, which):
        n = lib().pforacle_bpf_map_size(self.h, which)
        out = np.empty(max(n, 1), POINT_DTYPE)
        lib().pforacle_bpf_get_map(self.h, which, _vp(out))
        return out[:n].copy()

    def iter_poses(self):
        out = np.zeros((16, 7))
        n = lib().pforacle_bpf_iter_poses(self.h, _vp(out), 16)
        return out[:n].copy()

    def stats(self):
        out = np.zeros(8, np.int32)
        lib().pforacle_bpf_stats(self.h, _vp(out))
        return dict(zip(("n_beam_ds", "n_pillar_ds", "n_facade_ds", "n_line_res", "n_plane_res", "passes", "lm_iterations"), out.tolist()))
