"""Host-side multi-GPU logic on CPU: batch sharding and the shared-gradient all-reduce, world_size 2, gloo."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussigan_b200.dist import reduce_shared_grads, shard_batch, shard_range


def test_shard_range_partitions_the_batch():
    for B in (0, 1, 7, 64, 4096):
        for W in (1, 2, 3, 8):
            spans = [shard_range(B, W, r) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        B, K = 10, 6
        dmu = torch.randn(B, K, 1, 3, generator=g)
        dsig = torch.randn(B, K, 3, 3, generator=g, dtype=torch.float64)
        lmu, lsig = shard_batch([dmu, dsig], world, rank)
        rmu, rsig = reduce_shared_grads([lmu, lsig])
        ok = (torch.allclose(rmu, dmu.sum(0, keepdim=True), atol=1e-5)
              and torch.allclose(rsig, dsig.sum(0, keepdim=True), atol=1e-12)
              and rmu.dtype == torch.float32 and rsig.dtype == torch.float64 and rmu.shape == (1, K, 1, 3))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_reduce_shared_grads_world2_gloo():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_reduce_shared_grads_single_process():
    a = torch.arange(24, dtype=torch.float32).view(4, 2, 3)
    (r,) = reduce_shared_grads([a])
    assert torch.equal(r, a.sum(0, keepdim=True))
