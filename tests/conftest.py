import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _build_if_needed():
    so = os.path.join(ROOT, "master-thesis-lpf-in-mfem_b200", "liblpf_b200.so")
    oso = os.path.join(ROOT, "oracle", "liblpf_oracle.so")
    if not (os.path.exists(so) and os.path.exists(oso)):
        import __graft_entry__ as g
        g.build()


@pytest.fixture(scope="session")
def lpf():
    """The product package (C-ABI binding); aliased as lpf_b200."""
    _build_if_needed()
    mod = importlib.import_module("master-thesis-lpf-in-mfem_b200")
    sys.modules["lpf_b200"] = mod
    return mod


@pytest.fixture(scope="session")
def orc():
    """The numpy oracle (test infrastructure only)."""
    import lpf_oracle
    return lpf_oracle


@pytest.fixture(scope="session")
def corc():
    """The C oracle library (test infrastructure only)."""
    _build_if_needed()
    import oracle_c
    return oracle_c


@pytest.fixture(scope="session")
def cuda(lpf):
    import torch
    if not torch.cuda.is_available() or lpf.lib.lpf_device_count() < 1:
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    # the library enqueues on the stream it is given; a NULL stream means "create your own", so run the
    # whole session on an explicit non-default torch stream shared by torch ops and the C-ABI calls
    torch.cuda.set_stream(torch.cuda.Stream())
    return torch


MESH_DIR = os.path.join(ROOT, "tests", "meshes")
