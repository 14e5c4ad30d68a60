import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def pfb():
    from pf_loader import pfb as mod
    return mod


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    return O


@pytest.fixture(scope="session")
def capi(pfb):
    c = pfb.capi
    c.lib()
    if c.lib().pf_device_count() < 1:
        pytest.fail("no CUDA device visible: the product has no CPU fallback")
    return c


@pytest.fixture(scope="session")
def cfg2_scans(pfb):
    p = pfb.synth.config("cfg2")
    return p, [pfb.synth.scan(p, f) for f in range(4)]
