import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the product library (nvcc cross-compiles without a GPU) and the oracle (gcc)."""
    import splpak_b200
    from oracle import build_oracle

    splpak_b200.build()
    build_oracle()
    yield


@pytest.fixture(scope="session")
def oracle():
    from oracle import Oracle

    return Oracle()


@pytest.fixture(scope="session")
def oracle32():
    from oracle import Oracle

    return Oracle(real32=True)


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
