"""Shared helpers for the test-suite (golden loading, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SUB = 8
SIZES = (32, 64, 128, 256)

# ---- tolerances of the path (BASELINE.json north_star; SURVEY.md 7.2 for why they are norm-relative)
FWD_RTOL = 1e-5      # |out - ref| <= FWD_RTOL * max|ref|   (forward masks / maps)
GRAD_RTOL = 1e-4     # |g - ref|   <= GRAD_RTOL * max|ref|  (gradients)


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def inputs_of(gold, device="cpu"):
    return {k[3:]: torch.from_numpy(v).to(device) for k, v in gold.items() if k.startswith("in_")}


def weights_like(shapes, seed):
    g = torch.Generator().manual_seed(int(seed))
    return [torch.randn(tuple(s), generator=g, dtype=torch.float32) for s in shapes]


def sub(m):
    """The stored view of a map: full for n <= 64, every 8th pixel otherwise."""
    return m if m.shape[1] <= 64 else m[:, ::SUB, ::SUB, :]


def max_rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-300))


def assert_close_norm(a, b, rtol, what=""):
    e = max_rel(a, b)
    assert e <= rtol, f"{what}: max|diff|/max|ref| = {e:.3e} > {rtol:g}"
