import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _gpu_available() -> bool:
    try:
        from pdl_b200 import _abi
        return _abi.load().pdlb200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle_engine():
    from oracle_engine import OracleEngine
    return OracleEngine()


@pytest.fixture(scope="session")
def cuda_engine():
    from pdl_b200 import CudaEngine, set_default_engine
    if not _gpu_available():
        pytest.fail("-m gpu tests need a CUDA device and the built libpdlb200.so (no CPU fallback)")
    e = CudaEngine()
    set_default_engine(e)
    return e
