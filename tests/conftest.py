import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_lib
    return oracle_lib


@pytest.fixture(scope="session")
def ctx():
    from ambigram_b200 import capi
    c = capi.Context(device=0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx_small():
    """partitions of up to 16 segments through the warp-per-partition kernel first (AMBI_CFG_SMALL_KERNEL); whatever
    does not fit its record capacity continues on the level-synchronous global-memory path"""
    from ambigram_b200 import capi
    c = capi.Context(device=0, small_kernel=True)
    yield c
    c.close()


@pytest.fixture(params=["global", "small"])
def any_ctx(request, ctx, ctx_small):
    return ctx if request.param == "global" else ctx_small
