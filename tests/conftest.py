import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """libidash_b200.so, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    ge.build_native()
    from idash2019_2_b200 import _lib
    return _lib.lib()


@pytest.fixture(scope="session")
def gpu_ctx(built_lib):
    from idash2019_2_b200 import api
    ctx = api.Context(0)
    yield ctx
    ctx.close()
