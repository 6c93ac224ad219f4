import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
# the reference's own tests use paths relative to the repository root ("tests/data/...")
os.chdir(ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def data_dir() -> Path:
    return ROOT / "tests" / "data"


@pytest.fixture(scope="session")
def engine():
    """One ac_ctx for the whole GPU session; fails loudly when the CUDA library is missing."""
    from auriclass_b200.mash_gpu import MashGPU

    eng = MashGPU(0)
    yield eng
    eng.close()
