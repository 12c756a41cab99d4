import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# The parity tests drive the kernels through HOST buffers.  Host-buffer calls of deep-table chains normally run the
# "light" image (PCIe-bound calls, k_chain_mma.cu); TTN_MMA_LIGHT=0 makes them run the deep-table image instead, so
# that the kernel variant device-resident callers get (and bench.py's `value` times) is the one under test.
# tests/test_gpu_round2.py::test_light_variant_for_host_buffers covers the default.
os.environ.setdefault("TTN_MMA_LIGHT", "0")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    # built artefacts are git-ignored: on a fresh checkout compile them first (nvcc cross-compiles without a GPU)
    so = os.path.join(ROOT, "itensornumericalanalysis.jl_b200", "csrc", "libttneval.so")
    orc_so = os.path.join(ROOT, "oracle", "libttn_oracle.so")
    if not (os.path.exists(so) and os.path.exists(orc_so)):
        import __graft_entry__
        __graft_entry__.build()


def _have_gpu():
    from itna_b200 import _capi
    return _capi.lib().ttn_device_count() > 0  # raises if libttneval.so is missing


def pytest_collection_modifyitems(config, items):
    gpu_items = [i for i in items if "gpu" in i.keywords]
    if not gpu_items:
        return
    if not _have_gpu():
        skip = pytest.mark.skip(reason="no CUDA device on this host")
        for i in gpu_items:
            i.add_marker(skip)
