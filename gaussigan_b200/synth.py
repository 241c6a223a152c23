"""Synthetic renderer inputs drawn from the reference's activation ranges -- for benchmarks and
profiling drivers.  No oracle code: the covariances are built by the library's own ``sigma_from_frame`` kernel.

    mus   = tanh(N(0,1))                                   mask/mask_gen_dynamic.py:378
    Sigma = sigma_from_frame(tanh(N), tanh(N), N)          :383-401
    roty  = 180 * tanh(N(0,1)) degrees [+ U(0, angle)]     :458, :673       rotx = rotz = 0  (:706)
    target masks: random binary discs in {-1,+1}           stand-in for the dataset masks (:672)

Seed 56 is the reference's own (:21).  Tensors are generated on the CPU generator (reproducible) and returned on the
CPU, like the arrays a data loader would hand over; only the covariance construction runs on the GPU.
"""
from __future__ import annotations

import torch

from .geometry import sigma_from_frame


def synth_inputs(B: int, K: int, seed: int = 56, angle: float = 0.0, device="cuda"):
    """dict of CPU tensors: mus [B,K,1,3] f32, sigma [B,K,3,3] f64, rotx/roty/rotz [B,1] f32 (degrees)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    f32 = torch.float32
    mus = torch.tanh(torch.randn(B, K, 1, 3, generator=g, dtype=f32))
    v0 = torch.tanh(torch.randn(B, K, 3, generator=g, dtype=f32))
    v1 = torch.tanh(torch.randn(B, K, 3, generator=g, dtype=f32))
    u = torch.randn(B, K, 3, generator=g, dtype=f32)
    dev = torch.device(device)
    sigma = sigma_from_frame(v0.to(dev), v1.to(dev), u.to(dev)).cpu()
    roty = 180.0 * torch.tanh(torch.randn(B, 1, generator=g, dtype=f32))
    if angle:
        roty = roty + angle * torch.rand(B, 1, generator=g, dtype=f32)
    zr = torch.zeros_like(roty)
    return dict(mus=mus, sigma=sigma, rotx=zr.clone(), roty=roty, rotz=zr.clone())


def synth_target_masks(B: int, n: int, seed: int = 57) -> torch.Tensor:
    """Random binary discs in {-1,+1}, [B,n,n,1] f32."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    c = torch.rand(B, 2, generator=g) * 1.0 - 0.5
    r = torch.rand(B, 1, generator=g) * 0.4 + 0.2
    ax = torch.linspace(-1.0, 1.0, n)
    yy, xx = torch.meshgrid(ax, ax, indexing="ij")
    d2 = (xx[None] - c[:, 0, None, None]) ** 2 + (yy[None] - c[:, 1, None, None]) ** 2
    m = (d2 < (r[:, :, None] ** 2)).to(torch.float32) * 2.0 - 1.0
    return m.unsqueeze(-1)
