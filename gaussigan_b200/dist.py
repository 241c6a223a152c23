"""Multi-GPU plumbing for the renderer: one process per GPU, batch-sharded, no data-path collective.

Every batch element has its own Gaussians, camera, output tile and gradient, so the
renderer needs zero communication.  The only exchange is the training-time reduction of gradients of
parameters SHARED across the batch -- the canonical Gaussians (``mucann [1,K,1,3]``, ``sigmacann
[1,K,3,3]``, broadcast over the batch at mask/mask_gen_dynamic.py:701-702) -- which the reference
gets implicitly from in-graph tower replication (mask/GAN.py:121-151) and which here is one
``all_reduce(SUM)`` over NCCL/NVLink (``gloo`` in the CPU tests).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous batch slice [lo, hi) of rank ``rank``; sizes differ by at most one, earlier ranks larger."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    q, r = divmod(int(B), int(world_size))
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def shard_batch(tensors: Sequence[torch.Tensor], world_size: int, rank: int):
    """Slice every [B,...] tensor to this rank's shard."""
    lo, hi = shard_range(tensors[0].shape[0], world_size, rank)
    return [t[lo:hi] for t in tensors]


def reduce_shared_grads(per_sample_grads: Sequence[torch.Tensor], group=None) -> list:
    """Per-sample gradients of batch-broadcast parameters -> the gradient of the shared parameter.

    ``per_sample_grads[i]`` is ``[B_local, ...]``; returns ``[1, ...]`` tensors summed over the local batch
    and then over all ranks with a single flat ``all_reduce`` (one collective per step regardless of
    how many tensors there are -- the payload is 12*K numbers, so it is latency-bound by design).
    """
    local = [g.sum(dim=0, keepdim=True) for g in per_sample_grads]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    wide = torch.float64 if any(g.dtype == torch.float64 for g in local) else torch.float32
    flat = torch.cat([g.reshape(-1).to(wide) for g in local])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    out, off = [], 0
    for g in local:
        n = g.numel()
        out.append(flat[off:off + n].reshape(g.shape).to(g.dtype))
        off += n
    return out
