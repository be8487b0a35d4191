import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "late: run after every other test (see pytest_collection_modifyitems)")


def pytest_collection_modifyitems(config, items):
    """A GPU test that has not run on a B200 yet (written after a round's GPU budget was spent) can be marked `late`: it is ordered
    behind everything else, so that under `-x` a surprise in it cannot hide the tests that are known to pass.  None is marked now."""
    late = [it for it in items if it.get_closest_marker("late")]
    if late:
        items[:] = [it for it in items if not it.get_closest_marker("late")] + late


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def cuda_ctx():
    """One jl_ctx for the whole GPU test session; fails loudly if the CUDA library is missing."""
    from jlama_b200 import native
    ctx = native.Context(0)
    yield ctx
    ctx.close()


@pytest.fixture(scope="session")
def cuda_ops(cuda_ctx):
    from jlama_b200.ops import CudaTensorOperations
    return CudaTensorOperations(cuda_ctx)
