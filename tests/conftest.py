import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def port():
    """oracle/libsiril_oracle.so, built on demand from oracle/siril_oracle.c."""
    from oracle import port as p
    p.lib()
    return p


@pytest.fixture(scope="session")
def refrun():
    """oracle/_ref/libsiril_ref.so: the reference's own stacking.c, compiled unmodified.
    Built here when /root/reference exists, prebuilt (git-ignored, gpurun-shipped) otherwise."""
    from oracle import refrun as r
    if not r.available():
        import subprocess
        subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not r.available():
        pytest.skip("oracle/_ref/libsiril_ref.so not available")
    return r


@pytest.fixture(scope="session")
def pkg():
    """The product's host-side mirror (ctypes over libsiril_stack_cuda.so)."""
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    return g.load_package()


@pytest.fixture(scope="session")
def gpu(pkg):
    if pkg.device_count() < 1:
        pytest.fail("no CUDA device visible: the -m gpu tests must run on the GPU box")
    return pkg
