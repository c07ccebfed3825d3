import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu() -> bool:
    try:
        import ctypes
        import mppi_generic_b200 as m
        # cheap probe through the product library itself: creating an engine fails with NO_DEVICE on CPU boxes
        m.host.lib()
        e = m.Engine(m.CartpoleDynamics(), m.CartpoleQuadraticCost(), m.GaussianDistribution(1), 32, 4)
        e.close()
        return True
    except Exception:
        return False


_GPU = None


def pytest_collection_modifyitems(config, items):
    global _GPU
    if any("gpu" in it.keywords for it in items):
        if _GPU is None:
            _GPU = _has_gpu()
        if not _GPU:
            skip = pytest.mark.skip(reason="no CUDA device in this container (GPU tests run under gpurun)")
            for it in items:
                if "gpu" in it.keywords:
                    it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def mp():
    import mppi_generic_b200 as m
    return m
