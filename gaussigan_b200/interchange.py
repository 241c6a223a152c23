"""Interchange with the reference's on-disk / graph-boundary formats.

* the Gaussians the reference's inference script saves per input image (mask/infer_masks.py:447-448):
  ``m_%d.npy`` = ``mu3d [1,K,1,3]`` float32, ``sig3d_%d.npy`` = ``sigma3d [1,K,3,3]`` float64;
* the tensors fed into the frozen ``l_to_m`` decoder graph (mask/infer_masks.py:394-395, 455-456):
  ``gen/genlandmarks/mu2d [B,K,2]`` and ``gen/genlandmarks/sigma2d [B,K,2,2]`` (float32 in the inference script).

File I/O is host-side numpy; tensors land on the requested CUDA device.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np
import torch


def load_reference_gaussians(m_path: str, sig_path: str, device="cuda") -> Tuple[torch.Tensor, torch.Tensor]:
    """Read ``m_%d.npy`` / ``sig3d_%d.npy`` -> (mu3d [1,K,1,3] f32, sigma3d [1,K,3,3] f64) on ``device``."""
    mu = np.load(m_path)
    sig = np.load(sig_path)
    if sig.ndim != 4 or sig.shape[-2:] != (3, 3):
        raise ValueError(f"{sig_path}: expected [1,K,3,3], got {sig.shape}")
    K = sig.shape[1]
    if mu.size != sig.shape[0] * K * 3:
        raise ValueError(f"{m_path}: expected [1,K,1,3] with K={K}, got {mu.shape}")
    mu = torch.from_numpy(np.ascontiguousarray(mu, dtype=np.float32)).reshape(sig.shape[0], K, 1, 3)
    sig = torch.from_numpy(np.ascontiguousarray(sig, dtype=np.float64))
    return mu.to(device), sig.to(device)


def save_reference_gaussians(mu3d: torch.Tensor, sigma3d: torch.Tensor, directory: str, index: int) -> Tuple[str, str]:
    """Write the pair exactly as mask/infer_masks.py:447-448 does."""
    K = sigma3d.shape[-3]
    m_path = os.path.join(directory, "m_%d.npy" % index)
    s_path = os.path.join(directory, "sig3d_%d.npy" % index)
    np.save(m_path, mu3d.detach().cpu().reshape(-1, K, 1, 3).numpy().astype(np.float32))
    np.save(s_path, sigma3d.detach().cpu().reshape(-1, K, 3, 3).numpy().astype(np.float64))
    return m_path, s_path


def decoder_feed(mu2d: torch.Tensor, sigma2d: torch.Tensor) -> dict:
    """The feed_dict of the frozen decoder graph (mask/infer_masks.py:455-456), as host numpy float32."""
    return {"gen/genlandmarks/mu2d:0": mu2d.detach().float().cpu().numpy(),
            "gen/genlandmarks/sigma2d:0": sigma2d.detach().float().cpu().numpy()}
