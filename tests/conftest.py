import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def gpu_ctx():
    """Engine context on cuda:0.  Fails (does not skip) when the CUDA library is missing or unusable:
    the -m gpu tests must never pass on a fallback."""
    import sparse_linear_algebra_tests_b200 as pkg
    ctx = pkg.Context(0)
    pkg.set_default_context(ctx)
    yield ctx
    pkg.set_default_context(None)


@pytest.fixture
def cfg(gpu_ctx):
    """Change tuning switches of the engine context (b200_ctx_configure) for one test; restored afterwards."""
    saved = gpu_ctx.config()

    def set_(**fields):
        gpu_ctx.configure(**fields)

    yield set_
    gpu_ctx.restore(saved)
