"""pfilter-noetic_b200: B200-native (sm_100a) implementation of PFilter's per-frame LiDAR hot path.

The product is ``libpfilter_b200.so`` (hand-written CUDA behind the C ABI in ``include/pfilter_b200.h``).
This package is only the Python host-side harness over that C ABI (ctypes): ``capi`` binds the
``pf_*`` entry points, ``api`` mirrors the reference's class interface, ``synth`` generates the
synthetic sequences.  There is NO CPU fallback: if the CUDA library is missing, importing ``capi`` raises.
"""
from . import synth  # noqa: F401

__all__ = ["synth", "capi", "api"]


def __getattr__(name):
    if name in ("capi", "api"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
