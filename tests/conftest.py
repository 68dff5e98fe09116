import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import abwo
    abwo.build()
    return abwo


@pytest.fixture(autouse=True)
def _redzones_intact(request):
    """With ABW_REDZONE=1 in the environment (debug allocator of csrc/context.cu: canary zones around every device block, blocks handed out filled
    with 0xFF and never reused) every GPU test also asserts that no canary was damaged while it ran."""
    yield
    if os.environ.get("ABW_REDZONE", "0") not in ("", "0") and request.node.get_closest_marker("gpu") is not None:
        from abawaca_b200 import capi
        assert capi.load().abw_redzone_violations() == 0
