import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (runs the CUDA path through the C ABI)")


@pytest.fixture(scope="session")
def built_lib():
    """libqgpu.so, built in-tree (nvcc cross-compiles without a GPU)."""
    from query_engine_b200 import _ffi, build

    if not _ffi.lib_path().exists():
        build.build()
    return _ffi.load()


@pytest.fixture(scope="session")
def ctx(built_lib):
    """One device context for the whole GPU session.  Fails (does not skip) when no GPU is usable: the
    product has no CPU fallback and a `-m gpu` run on a box without a GPU must be red."""
    from query_engine_b200 import _ffi

    c = _ffi.Context(int(os.environ.get("QGPU_TEST_DEVICE", "0")))
    yield c
    c.close()


@pytest.fixture(params=["interp", "jit"])
def executor(ctx, request):
    """The same operator tests run through both forms of the scan kernels: the ahead-of-time interpreter
    ("scan.jit" = 0) and the NVRTC-specialised kernel of the same body ("scan.jit" = 1)."""
    from query_engine_b200.executor import QueryExecutor

    ctx.set_option("scan.jit", 0 if request.param == "interp" else 1)
    yield QueryExecutor(ctx=ctx)
    ctx.set_option("scan.jit", 2)
