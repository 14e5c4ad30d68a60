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
