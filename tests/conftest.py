import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
# the reference's own tests use paths relative to the repository root ("tests/data/...")
os.chdir(ROOT)
# Tests that emulate several ranks of the peer-memory exchange inside ONE process on one GPU keep a kernel of "rank" A
# spinning on a flag that a kernel of "rank" B sets.  CUDA multiplexes streams onto a few hardware queues (8 by default)
# and two streams that share a queue run in order -- B's kernel would sit behind A's spin until it times out.  One queue
# per stream (read when the CUDA context is created, so it has to be set before anything touches the GPU).  Real ranks
# are separate processes on separate GPUs and need nothing of the kind.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")


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
