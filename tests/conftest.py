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
# spinning on a flag that a kernel of "rank" B sets.  Two things of the CUDA runtime then matter, both read when the
# context is created (so they are set here, before anything touches the GPU):
#  * lazy module loading (the default) loads a kernel at its first launch, and that load waits for running kernels --
#    rank B's first K1 / K2 launch would sit behind rank A's spin until the spin times out (measured: 20 s per wait;
#    with EAGER loading the same test runs in milliseconds).  NVIDIA's guidance for kernels that must run concurrently;
#  * streams are multiplexed onto 8 hardware queues by default and two streams on one queue run in order.
# Real ranks are separate processes on separate GPUs: a rank's wait only ever depends on kernels its peers launched
# BEFORE their own wait, so neither setting is needed there.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")
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
