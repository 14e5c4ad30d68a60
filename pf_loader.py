"""Imports the package that lives in the (non-identifier) directory ``pfilter-noetic_b200/``.

    from pf_loader import pfb      # -> module object of pfilter-noetic_b200/__init__.py
"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG_DIR = os.path.join(ROOT, "pfilter-noetic_b200")
_NAME = "pfilter_noetic_b200"


def load():
    if _NAME in sys.modules:
        return sys.modules[_NAME]
    spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[_NAME] = mod
    spec.loader.exec_module(mod)
    return mod


pfb = load()
