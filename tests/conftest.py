import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle20():
    from oracle.oracle import Oracle
    return Oracle(20, 4)


@pytest.fixture(scope="session")
def oracle7():
    from oracle.oracle import Oracle
    return Oracle(7, 2)


@pytest.fixture(scope="session")
def engine20():
    import torch
    from blokus_rl_b200 import BlokusEngine
    assert torch.cuda.is_available(), "GPU tests need a B200"
    return BlokusEngine(20, 4)


@pytest.fixture(scope="session")
def engine7():
    from blokus_rl_b200 import BlokusEngine
    return BlokusEngine(7, 2)
