import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The plain-C restatement (oracle/liboracle.so); always available after __graft_entry__.build()."""
    from oracle.binding import Oracle, available
    if not available("port"):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return Oracle("port")


@pytest.fixture(scope="session")
def ref():
    """The reference's own sources (oracle/_ref/libdmz_ref.so).  Only exists where /root/reference was
    available at build time (it travels to the GPU box as a prebuilt, git-ignored file)."""
    from oracle.binding import Oracle, available
    if not available("ref"):
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return Oracle("ref")


@pytest.fixture(scope="session")
def refx():
    """The reference's sources built with SCAN_EXPIRY=1 (oracle/_ref/libdmz_ref_expiry.so): the expiry taps."""
    from oracle.binding import Oracle, available
    if not available("refx"):
        pytest.skip("oracle/_ref/libdmz_ref_expiry.so not built (no /root/reference on this machine)")
    return Oracle("refx")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(HERE, "golden", "ref_golden.npz"))


@pytest.fixture(scope="session")
def golden_formats():
    """dmz_YCbCr_to_RGB / dmz_deinterleave_RGBA_to_R / Cython stencil outputs of the reference build (tools/make_ref_formats_golden.py)."""
    return np.load(os.path.join(HERE, "golden", "ref_formats.npz"))


@pytest.fixture(scope="session")
def pkg():
    from util import load_pkg
    return load_pkg()


@pytest.fixture(scope="session")
def dmz(pkg):
    """A GPU context.  Deliberately NOT skipped when creation fails: on a GPU box a missing / broken CUDA
    extension must fail the gpu-marked tests loudly."""
    d = pkg.Dmz(device=0)
    yield d
    d.close()
