"""Real multi-process run of the row-sharded path (torchrun, one rank per GPU, CUDA-IPC peer
mailboxes / NCCL) against the single-GPU engine: tools/mp_parity.py.  Needs >= 2 GPUs (skipped on
a one-GPU box; run there by `gpurun --gpus 2 -- bash tools/mp_parity.sh 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
@pytest.mark.parametrize("storage", ["lower", "full"])
def test_two_process_parity(cuda, storage, exchange):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, TAL_EXCHANGE=exchange)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541",
                        os.path.join(ROOT, "tools", "mp_parity.py"), "small", storage, "incremental"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mp parity ok" in r.stdout


@pytest.mark.parametrize("q", [1, 2])
def test_two_process_fused_step(cuda, q):
    """The fused step (tal_ctx_step: records published by the finish kernel, combined in the push kernel)
    on two ranks over CUDA-IPC peer memory against the single-GPU fused run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, TAL_EXCHANGE="peer")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29542",
                        os.path.join(ROOT, "tools", "mp_parity.py"), "fused", "lower", str(q)],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mp parity ok" in r.stdout
