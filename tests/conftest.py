"""pytest configuration: registers the ``gpu`` marker and puts the repo root on ``sys.path``.

CPU suite (driver, every round):   python -m pytest tests/ -x -q -m "not gpu"
GPU suite (B200, round end):       python -m pytest tests/ -x -q -m gpu
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests"), str(ROOT / "tests" / "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
