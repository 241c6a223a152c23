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


class P2PSharedGradReducer:
    """``reduce_shared_grads`` for the renderer's own gradients over NVLink peer memory (include/gaussigan.h):
    ``post`` = local batch sum + stores into every peer's receive slot + flags, ``wait`` = flag wait + rank-ordered sum.
    An order of magnitude less latency than an eager NCCL ``all_reduce`` at this payload (12*K numbers), bit-identical
    on all ranks, replayable from a CUDA graph, and the two halves can be separated so the exchange overlaps other work.

        red = P2PSharedGradReducer(K, device)                  # collective: every rank of the group constructs it
        dmu_sum, dsig_sum = red(dmus, dsigma)                  # [B,K,3] f32, [B,K,3,3] f64 -> [K,3], [K,3,3] float64
        red.post(dmus, dsigma); ...; dmu_sum, dsig_sum = red.wait()      # the same, overlapped

    Needs ``torch.distributed`` initialised with NCCL on one node with peer access (NVLink / NVSwitch): the exchange
    buffers come from ``torch.distributed._symmetric_memory``.  ``check()`` synchronises and raises if a peer timed out.
    """

    def __init__(self, K: int, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from ._lib import lib
        dev = torch.device(device)
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.K, self.dev = int(K), dev
        group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        nbytes = int(lib.gg_p2p_buffer_bytes(self.K, self.world))
        self.buf = symm_mem.empty((nbytes + 7) // 8, dtype=torch.int64, device=dev)
        self.handle = symm_mem.rendezvous(self.buf, group)
        self.buf.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group)                                      # nobody signals into a buffer that is still being zeroed
        self.peers = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=dev)
        self.out = torch.empty(12 * self.K, dtype=torch.float64, device=dev)

    def post(self, dmus: torch.Tensor, dsigma: torch.Tensor) -> None:
        """First half: local batch sum, stores into every peer's slot, flags (``gg_shared_grad_post_p2p``).  Enqueue it
        right behind the backward; at most one post may be outstanding (``wait()`` before the next ``post()``)."""
        from ._lib import check, lib
        B = dmus.shape[0]
        assert dmus.is_cuda and dmus.dtype == torch.float32 and dmus.is_contiguous() and dmus.numel() == B * self.K * 3
        assert dsigma.dtype == torch.float64 and dsigma.is_contiguous() and dsigma.numel() == B * self.K * 9
        st = torch.cuda.current_stream(self.dev).cuda_stream
        check(lib.gg_shared_grad_post_p2p(dmus.data_ptr(), dsigma.data_ptr(), B, self.K, self.peers.data_ptr(),
                                          self.world, self.rank, self.dev.index, st), "gg_shared_grad_post_p2p")

    def wait(self):
        """Second half: wait for the peers' flags and add the slots in rank order (``gg_shared_grad_wait_p2p``), on the
        current stream (which must be ordered after the post).  Returns ([K,3], [K,3,3]) float64 views of ``self.out``;
        they hold NaN if a peer did not arrive within ``GG_P2P_TIMEOUT_MS``."""
        from ._lib import check, lib
        st = torch.cuda.current_stream(self.dev).cuda_stream
        check(lib.gg_shared_grad_wait_p2p(self.K, self.peers.data_ptr(), self.world, self.rank, self.out.data_ptr(),
                                          self.dev.index, st), "gg_shared_grad_wait_p2p")
        K = self.K
        return self.out[:3 * K].view(K, 3), self.out[3 * K:].view(K, 3, 3)

    def __call__(self, dmus: torch.Tensor, dsigma: torch.Tensor):
        self.post(dmus, dsigma)
        return self.wait()

    def check(self) -> None:
        """Synchronise and raise if any wait timed out (the result of that call is NaN as well)."""
        torch.cuda.synchronize(self.dev)
        err = int(self.buf.view(torch.int32)[1].item())
        if err:
            raise RuntimeError(f"gg_shared_grad_wait_p2p: a peer did not arrive in call {err}; the results of that call "
                               "are NaN and the reducer must be rebuilt")
