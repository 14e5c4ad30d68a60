"""ctypes binding of the synthetic LiDAR generator (csrc/synth.cpp, include/pf_synth.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SCENE_STREET, SCENE_CAMPUS, SCENE_CAMPUS_DENSE, SCENE_STREET_DENSE = 0, 1, 2, 3
TRAJ_STREET, TRAJ_LOOP = 0, 1


class SynthParams(C.Structure):
    _fields_ = [("sensor_lines", C.c_int32), ("azimuth_steps", C.c_int32), ("seed", C.c_uint64),
                ("scene", C.c_int32), ("trajectory", C.c_int32), ("speed", C.c_double),
                ("range_sigma", C.c_double), ("elev_jitter_deg", C.c_double),
                ("min_range", C.c_double), ("max_range", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libpf_synth.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(path)
        _lib.pf_synth_default_params.argtypes = [C.POINTER(SynthParams)]
        _lib.pf_synth_pose.argtypes = [C.POINTER(SynthParams), C.c_int, C.POINTER(C.c_double)]
        _lib.pf_synth_scan.argtypes = [C.POINTER(SynthParams), C.c_int, C.c_void_p, C.c_int]
        _lib.pf_synth_scan.restype = C.c_int
    return _lib


def params(**kw):
    p = SynthParams()
    lib().pf_synth_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise KeyError(k)
        setattr(p, k, v)
    return p


def scan(p, frame):
    """Returns an (n, 4) float32 array x,y,z,intensity in the sensor frame."""
    cap = p.sensor_lines * p.azimuth_steps
    buf = np.empty((cap, 4), dtype=np.float32)
    n = lib().pf_synth_scan(C.byref(p), frame, buf.ctypes.data, cap)
    if n < 0:
        raise RuntimeError("pf_synth_scan overflow")
    return np.ascontiguousarray(buf[:n])


def pose(p, frame):
    """Ground-truth pose [qx qy qz qw tx ty tz] (float64)."""
    out = (C.c_double * 7)()
    lib().pf_synth_pose(C.byref(p), frame, out)
    return np.array(out[:], dtype=np.float64)


# Named sequence configurations of BASELINE.json
def config(name):
    if name in ("cfg1", "cfg2", "cfg3"):     # 64-ring street, 1 m/frame
        return params(sensor_lines=64, seed=2022)
    if name == "cfg4":                       # 32-ring slow campus loop
        return params(sensor_lines=32, seed=2023, scene=SCENE_CAMPUS, trajectory=TRAJ_LOOP, speed=0.15)
    if name == "cfg4s":                      # 32-ring, low speed, along the planes+poles street (tracks for thousands of frames)
        return params(sensor_lines=32, seed=2023, speed=0.15)
    if name == "cfg4sd":                     # the same with volumetric scatter beside the lane: map-growth stress (filter off)
        return params(sensor_lines=32, seed=2023, scene=SCENE_STREET_DENSE, speed=0.15)
    if name == "cfg4d":                      # the same loop through the campus with volumetric scatter (foliage): map-growth stress
        return params(sensor_lines=32, seed=2023, scene=SCENE_CAMPUS_DENSE, trajectory=TRAJ_LOOP, speed=0.15)
    if name.startswith("cfg5."):             # 8 independent 64-ring sequences, seeds 3000..3007
        return params(sensor_lines=64, seed=3000 + int(name.split(".")[1]))
    raise KeyError(name)
