"""Multi-rank GPU test of the peer-memory exchange: two ranks under torchrun -- values against NCCL,
bit-identity across ranks, a rank delayed by 1.5 s, the post/wait split, and a peer that times out (NaN results +
latched error instead of stale sums)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(extra_args, env_extra, port):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "p2p_allreduce_check.py")] + extra_args
    return subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)


@pytest.fixture(scope="module")
def two_gpus():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs with peer access (run with gpurun --gpus 2)")


def test_p2p_exchange_two_ranks(two_gpus):
    r = _torchrun([], {}, 29631)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for needle in ("20 calls match NCCL and are bit-identical across ranks: True",
                   "delayed by 1.5 s still yields the correct sum everywhere: True",
                   "post / wait with work in between: True"):
        assert needle in r.stdout, r.stdout[-3000:]


def test_p2p_exchange_timeout_poisons_the_result(two_gpus):
    r = _torchrun(["--timeout-case"], {"GG_P2P_TIMEOUT_MS": "500"}, 29633)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "timed-out peer -> NaN results and a latched error on the waiting ranks: True" in r.stdout
