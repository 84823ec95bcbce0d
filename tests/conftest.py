import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import b2oracle
    b2oracle.lib()
    return b2oracle


@pytest.fixture(scope="session")
def z():
    """The product package; the library must already be built (no silent fallback)."""
    import zig_lz4_b200
    zig_lz4_b200.lib()
    # B2_TEST_TUNE="k1_variant=2,k2_variant=1": run the same parity tests against an experimental kernel variant
    # (the knob is the library's b2lz4_debug_tune; the library itself never reads the environment)
    for kv in filter(None, os.environ.get("B2_TEST_TUNE", "").split(",")):
        k, v = kv.split("=")
        zig_lz4_b200.debug_tune(k, int(v))
    return zig_lz4_b200


@pytest.fixture(scope="session")
def ctx(z):
    c = z.Context(0)
    yield c
    c.close()
