import os
import sys

# The single-GPU emulation of the peer-memory protocol (tests/test_gpu_sharded.py, exchange="peer") runs
# shards whose kernels WAIT ON EACH OTHER inside one CUDA context.  With lazy module loading the first
# launch of a kernel in one thread loads its module under a context-wide lock while the other thread's
# kernel spins, which stalls until that kernel's time-out; load everything up front instead.  (Between
# processes on separate GPUs — the real deployment — nothing is shared and this cannot occur.)
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "transductive-active-learning_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")


@pytest.fixture(scope="session")
def lib_path():
    import importlib.util
    spec = importlib.util.spec_from_file_location("tal_build", os.path.join(PKG, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


@pytest.fixture(scope="session")
def cuda(lib_path):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda", 0)
