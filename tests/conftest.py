import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rust-tracer_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _cuda_devices():
    try:
        import rtrace_b200 as rt
        return rt.device_count()
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must FAIL loudly (no silent fallback); a
    # plain `pytest tests/` on a CPU box skips the GPU tests instead.
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    _oracle.lib()
    return _oracle


@pytest.fixture(scope="session")
def rt():
    import rtrace_b200
    rtrace_b200.lib()
    return rtrace_b200


@pytest.fixture(scope="session")
def oracle_scene8(oracle):
    return oracle.Scene()


@pytest.fixture(scope="session")
def gpu_scene8(rt):
    return rt.Scene()
