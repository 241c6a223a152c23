"""TEST INFRASTRUCTURE ONLY -- float64 torch-CPU restatement of GaussiGAN's renderer.

Op-for-op restatement (same operation order, same dtypes at every cast
boundary) of the reference's splat-to-silhouette path, so that autograd through
it yields the reference's gradients.  Every function cites the reference
``file:line`` it follows (paths relative to ``/root/reference``).

PARITY UNPINNED against a running TensorFlow 1.12 (not installable here); pinned
against the reference's own source executed through ``oracle/tf_shim.py`` (see
``tests/golden/make_golden.py`` and ``tests/test_oracle.py``).

Not imported by ``gaussigan_b200``; see ``oracle/__init__.py``.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

F32 = torch.float32
F64 = torch.float64

#: the pyramid the reference hard-codes: ``2 ** (8 - 4 + i + 1)``, i = 0..3
#: (mask/mask_gen_dynamic.py:236, 315-316)
REFERENCE_SIZES = (32, 64, 128, 256)


# --------------------------------------------------------------------------- grid
def tf_linspace_f32(n: int) -> np.ndarray:
    """``tf.linspace(-1.0, 1.0, n)`` as TensorFlow 1.12 evaluates it, in float32.

    TF 1.12's LinSpace kernel computes ``step = (stop - start) / (num - 1)`` and
    ``out[i] = start + step * i`` in the op's dtype (float32 here:
    mask/mask_gen_dynamic.py:121-122).  ``torch.linspace`` uses a different
    (symmetric) formula, so the grid is rebuilt explicitly.  n == 1 yields
    ``[start]``.
    """
    start = np.float32(-1.0)
    stop = np.float32(1.0)
    if n == 1:
        return np.array([start], dtype=np.float32)
    step = np.float32((stop - start) / np.float32(n - 1))
    idx = np.arange(n, dtype=np.float32)
    return (start + (step * idx).astype(np.float32)).astype(np.float32)


# ------------------------------------------------------------------------ raster
def get_gaussian_maps_2d(mu: torch.Tensor, sigma: torch.Tensor, shape_hw: Sequence[int],
                         raster_dtype: torch.dtype = F64) -> torch.Tensor:
    """mask/mask_gen_dynamic.py:112-143 (float64) and mask/infer_masks.py:345-373 (float32).

    mu [B,K,2] (x = image column first, y = row second), sigma [B,K,2,2]
    -> [B,H,W,K] in ``raster_dtype``.  ``g = exp(-(X P X^T))`` with
    ``P = inv(sigma)`` (not symmetrised), no 1/2 factor, no normalisation.
    """
    H, W = int(shape_hw[0]), int(shape_hw[1])
    K = mu.shape[-2]
    # :121-122 -- float32 linspace, then up-cast (training) / kept f32 (inference)
    y = torch.from_numpy(tf_linspace_f32(H)).to(raster_dtype)
    x = torch.from_numpy(tf_linspace_f32(W)).to(raster_dtype)
    # :124-127 -- meshgrid('xy'): xg[i, j] = x[j], yg[i, j] = y[i]
    xg = x.view(1, W).expand(H, W)
    yg = y.view(H, 1).expand(H, W)
    xy = torch.stack([xg, yg], dim=-1)                      # [H,W,2]
    xy = xy.view(1, 1, H, W, 2).expand(1, K, H, W, 2)
    mu_ = mu.reshape(-1, K, 1, 1, 2)                        # :128
    inv = torch.linalg.inv(sigma)                           # :131
    inv = inv.to(raster_dtype).reshape(-1, K, 1, 2, 2)      # :133 (cast is infer_masks.py:363)
    pp = inv.expand(-1, K, W, 2, 2)                         # :134 tile over axis 2 (needs H == W)
    X = xy - mu_                                            # :135  [B,K,H,W,2]
    dist = torch.matmul(X, pp)                              # :136  batch dims [B,K,H] x [B,K,W]
    dist = (dist * X).sum(dim=-1)                           # :137
    g = torch.exp(-dist)                                    # :139
    return g.permute(0, 2, 3, 1)                            # :141  -> [B,H,W,K]


# -------------------------------------------------------------------- projection
def cos_f32(a: torch.Tensor) -> torch.Tensor:
    """float32 cosine, CORRECTLY ROUNDED: round_f32(cos_f64(a)).

    The reference evaluates ``tf.cos`` on float32 (mask/mask_gen_dynamic.py:246-252);
    Eigen's ``pcos``, glibc ``cosf``, Sleef and CUDA ``cosf`` all differ from one another
    by up to 1 ulp, and 1 ulp on R moves ``mu2d`` by ~1e-6 (measured), i.e. it is the
    dominant parity-noise term.  The oracle therefore pins the primitive to the
    correctly rounded value, which every one of those implementations approximates;
    the CUDA kernel evaluates the same definition.  TF's own last-ulp behaviour stays
    unpinned (<= 1 ulp of float32)."""
    return torch.cos(a.to(F64)).to(a.dtype) if a.dtype == F32 else torch.cos(a)


def sin_f32(a: torch.Tensor) -> torch.Tensor:
    """float32 sine, correctly rounded (see :func:`cos_f32`)."""
    return torch.sin(a.to(F64)).to(a.dtype) if a.dtype == F32 else torch.sin(a)


def matmul3_f32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """3x3 float32 matmul with a pinned evaluation order: ((a_i0*b_0j + a_i1*b_1j) + a_i2*b_2j),
    each product and sum rounded to float32 (no FMA).  ``tf.matmul`` on float32
    (mask/mask_gen_dynamic.py:255) leaves the order to Eigen; pinning it removes a
    1-ulp ambiguity that only arises when two of the three Euler angles are non-zero
    (every call site passes rotx = rotz = 0, making R = RY exactly)."""
    prods = a.unsqueeze(-1) * b.unsqueeze(-3)            # [...,i,k,j]
    return (prods[..., 0, :] + prods[..., 1, :]) + prods[..., 2, :]


def _euler_rotation_f32(rotx: torch.Tensor, roty: torch.Tensor, rotz: torch.Tensor) -> torch.Tensor:
    """mask/mask_gen_dynamic.py:240-255.  rot* [B,1] float32 DEGREES -> R [B,1,3,3] float32.

    ``tf.stack([tf.stack(col, -1), ...], -1)``: the inner lists are matrix
    COLUMNS, so  RX=[[1,0,0],[0,c,s],[0,-s,c]], RY=[[c,0,-s],[0,1,0],[s,0,c]],
    RZ=[[c,s,0],[-s,c,0],[0,0,1]];  R = RX @ RY @ RZ, all in float32.
    """
    pi32 = torch.tensor(np.pi, dtype=F32)
    d180 = torch.tensor(180.0, dtype=F32)
    ax = rotx.to(F32) * pi32 / d180                          # :240-242 (f32 mul, then f32 div)
    ay = roty.to(F32) * pi32 / d180
    az = rotz.to(F32) * pi32 / d180
    zr = torch.zeros_like(ay)
    on = torch.ones_like(ay)

    def cols(c0, c1, c2):
        return torch.stack([torch.stack(c0, -1), torch.stack(c1, -1), torch.stack(c2, -1)], -1)

    RX = cols([on, zr, zr], [zr, cos_f32(ax), -sin_f32(ax)], [zr, sin_f32(ax), cos_f32(ax)])
    RY = cols([cos_f32(ay), zr, sin_f32(ay)], [zr, on, zr], [-sin_f32(ay), zr, cos_f32(ay)])
    RZ = cols([cos_f32(az), -sin_f32(az), zr], [sin_f32(az), cos_f32(az), zr], [zr, zr, on])
    return matmul3_f32(matmul3_f32(RX, RY), RZ)              # :255


def project(mus: torch.Tensor, sigma: torch.Tensor, rotx: torch.Tensor, roty: torch.Tensor,
            rotz: torch.Tensor, tx: float, ty: float, tz: float, focal: float = 1.0
            ) -> Tuple[torch.Tensor, torch.Tensor]:
    """mask/mask_gen_dynamic.py:258-309: 3-D Gaussians + camera -> (mu2d [B,K,2], sigma2d [B,K,2,2]) f64.

    mus [B,K,1,3] f32, sigma [B,K,3,3] f64, rot* [B,1] f32 degrees.
    """
    B, K = mus.shape[0], mus.shape[1]
    R = _euler_rotation_f32(rotx, roty, rotz)                                   # [B,1,3,3] f32
    t = torch.tensor([[tx, ty, tz]], dtype=F64).view(1, 1, 1, 3).expand(1, K, 1, 3)  # :258-260
    # :264-272 intrinsics; inner lists are columns -> Kmat = [[f,0,px],[0,f,py],[0,0,1]], px = py = 0
    Kmat = torch.zeros(B, K, 3, 3, dtype=F64)
    Kmat[..., 0, 0] = float(np.float32(focal))
    Kmat[..., 1, 1] = float(np.float32(focal))
    Kmat[..., 2, 2] = 1.0

    sigma = sigma.to(F64)
    R = R.to(F64) * torch.ones_like(sigma)                                      # :274
    sig_c = torch.matmul(torch.matmul(R, sigma), R.transpose(-1, -2))           # :275
    inv = torch.linalg.inv(sig_c)                                               # :276
    m = mus.to(F64)                                                             # :277
    m = torch.matmul(R, m.transpose(-1, -2)).transpose(-1, -2) + t              # :278  [B,K,1,3]

    M0 = torch.matmul(inv, torch.matmul(m.transpose(-1, -2), m))                # :280
    M0 = torch.matmul(M0, inv.transpose(-1, -2))                                # :281
    M1 = torch.matmul(torch.matmul(m, inv), m.transpose(-1, -2)) - 1.0          # :282  [B,K,1,1]
    M1 = M1 * inv                                                               # :283
    M = M0 - M1                                                                 # :285
    flip = torch.tensor([[1., 1., 0.], [1., 1., 0.], [0., 0., 1.]], dtype=F64).view(1, 1, 3, 3)
    M = -M + 2 * M * flip                                                       # :287-289

    def minor(rows, cols_):
        return M[:, :, rows, :][:, :, :, cols_]
    M33 = minor([0, 1], [0, 1])                                                 # :290
    K33 = Kmat[:, :, [0, 1], :][:, :, :, [0, 1]]                                # :291
    M31 = minor([0, 1], [1, 2])                                                 # :292
    M23 = minor([0, 2], [0, 1])                                                 # :293
    det31 = torch.linalg.det(M31)                                               # :294
    det23 = torch.linalg.det(M23)
    det33 = torch.linalg.det(M33)
    detM = torch.linalg.det(M)                                                  # :297

    kmul = torch.stack([det31, -det23], dim=-1).reshape(-1, K, 2, 1)            # :299-300
    mup0 = torch.matmul(K33, kmul).squeeze(-1) / det33.reshape(-1, K, 1)        # :301
    mup1 = torch.stack([Kmat[:, :, 0, 2], Kmat[:, :, 1, 2]], dim=-1)            # :302
    mu2d = mup0 + mup1                                                          # :303
    sigma_w = (detM / det33).reshape(-1, K, 1, 1)                               # :305-307
    sigma2d = -sigma_w * torch.linalg.inv(M33)                                  # :308-309
    return mu2d, sigma2d


def get_landmarks(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1.0,
                  sizes: Sequence[int] = REFERENCE_SIZES) -> List[torch.Tensor]:
    """mask/mask_gen_dynamic.py:233-323 -> list of [B,n,n,K] float32 maps (n in ``sizes``)."""
    assert mus is not None and sigma is not None                                # :234-235
    mu2d, sigma2d = project(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal)
    return [get_gaussian_maps_2d(mu2d, sigma2d, [n, n]).to(F32) for n in sizes]  # :315-319


def get_landmarks_infer(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1.0, size: int = 256):
    """mask/infer_masks.py:186-267: f64 projection, cast to f32, f32 raster at 256 only.

    Returns (mu2d f32 [B,K,2], sigma2d f32 [B,K,2,2], maps f32 [B,size,size,K]).
    """
    mu2d, sigma2d = project(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal)
    mu2d = mu2d.to(F32)                                                         # :261
    sigma2d = sigma2d.to(F32)                                                   # :262
    maps = get_gaussian_maps_2d(mu2d, sigma2d, [size, size], raster_dtype=F32)  # :265, :345-373
    return mu2d, sigma2d, maps


# -------------------------------------------------------------------- consumers
def sum_composite(maps: torch.Tensor) -> torch.Tensor:
    """mask/mask_gen_dynamic.py:785-786: unclamped sum over K, keepdims -> [B,H,W,1]."""
    return maps.sum(dim=-1, keepdim=True)


def mask_recon_loss_bis(mask_pm1: torch.Tensor, maps256: torch.Tensor) -> torch.Tensor:
    """mask/mask_gen_dynamic.py:785-787: mean | (mask+1)/2 - sum_k g_k |, mask in [-1,1], [B,H,W,1]."""
    return ((mask_pm1 + 1) * 0.5 - sum_composite(maps256)).abs().mean()


def gm_cycle_loss(maps_a: torch.Tensor, maps_b: torch.Tensor) -> torch.Tensor:
    """mask/mask_gen_dynamic.py:791: mean |g_pose - g_cyc| over [B,256,256,K]."""
    return (maps_a - maps_b).abs().mean()


def colorize_landmark_maps(maps: torch.Tensor, colors: torch.Tensor) -> torch.Tensor:
    """mask/mask_gen_dynamic.py:221-231 / mask/infer_masks.py:174-184.

    maps [B,H,W,K], colors [K,3] -> [B,H,W,3] = max_k maps[...,k] * colors[k].
    """
    hm = maps.unsqueeze(-1) * colors.view(1, 1, 1, -1, 3).to(maps.dtype)
    return hm.max(dim=3).values


# ----------------------------------------------------- Gaussian construction (a6)
def l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """``tf.math.l2_normalize``: x * rsqrt(max(sum(x^2), eps))."""
    return x * torch.rsqrt(torch.clamp((x * x).sum(dim=-1, keepdim=True), min=eps))


def sigma_from_frame(v0_raw: torch.Tensor, v1_raw: torch.Tensor, u_raw: torch.Tensor) -> torch.Tensor:
    """mask/mask_gen_dynamic.py:383-401 (tail of get_musigma).

    v0_raw, v1_raw = tanh(dense) [N,K,3] f32; u_raw = dense pre-activation [N,K,3] f32
    -> Sigma [N,K,3,3] float64 (built in float32, then cast: :400-401).
    """
    v0 = l2_normalize(v0_raw.to(F32))                                           # :388
    v1 = torch.linalg.cross(v0, v1_raw.to(F32), dim=-1)                         # :389
    v1 = l2_normalize(v1)                                                       # :390
    v2 = torch.linalg.cross(v1, v0, dim=-1)                                     # :391
    V = torch.stack([v0, v2, v1], dim=-1)                                       # :392-396 columns v0,v2,v1
    u = torch.sigmoid(u_raw.to(F32)) * 0.5 + 1e-2                               # :397
    U = torch.diag_embed(u)                                                     # :399
    sig = torch.matmul(torch.matmul(V, U), V.transpose(-1, -2))                 # :400 (f32)
    return sig.to(F64)                                                          # :401


# ------------------------------------------------------------- deformation (a7)
def _euler_rotation_rad_f32(ax, ay, az):
    """mask/mask_gen_dynamic.py:442-453: same column convention as get_landmarks, RADIANS, per Gaussian."""
    zr = torch.zeros_like(az)
    on = torch.ones_like(az)

    def cols(c0, c1, c2):
        return torch.stack([torch.stack(c0, -1), torch.stack(c1, -1), torch.stack(c2, -1)], -1)

    RX = cols([on, zr, zr], [zr, cos_f32(ax), -sin_f32(ax)], [zr, sin_f32(ax), cos_f32(ax)])
    RY = cols([cos_f32(ay), zr, sin_f32(ay)], [zr, on, zr], [-sin_f32(ay), zr, cos_f32(ay)])
    RZ = cols([cos_f32(az), -sin_f32(az), zr], [sin_f32(az), cos_f32(az), zr], [zr, zr, on])
    return matmul3_f32(matmul3_f32(RX, RY), RZ)


def deform(mu_t_raw, scale_raw, rx_raw, ry_raw, rz_raw, theta_raw, cannmu, cannsigma):
    """mask/mask_gen_dynamic.py:426-458 (tail of dynamic transform_mu_sigma).

    Inputs are the dense heads AFTER tanh:  mu_t_raw [B,K*3], scale_raw [B,K*3],
    r*_raw [B,K], theta_raw [B,1] (all f32 in (-1,1));  cannmu [1,K,1,3] f32,
    cannsigma [1,K,3,3] f64.
    Returns mus [B,K,1,3] f32, sigmas [B,K,3,3] f64, theta [B,1] f32 (degrees).
    """
    K = cannmu.shape[1]
    cannsigma = cannsigma.to(F64)                                               # :409
    mu_t = (mu_t_raw.to(F32) * 0.1).reshape(-1, K, 1, 3)                        # :426-427
    mus = cannmu * torch.ones_like(mu_t) + mu_t                                 # :428
    U, S, _ = torch.linalg.svd(cannsigma, full_matrices=True)                   # :430 (tf returns S,U,V)
    Sd = torch.diag_embed(torch.sqrt(S))                                        # :431
    sc = (scale_raw.to(F32) * 0.05 + 1.0).reshape(-1, K, 3)                     # :432-433
    sc = torch.diag_embed(sc).to(F64)                                           # :435-436
    Ssc = Sd * sc                                                               # :437 elementwise
    ax = rx_raw.to(F32) * np.pi / 10.0                                          # :439-441
    ay = ry_raw.to(F32) * np.pi / 10.0
    az = rz_raw.to(F32) * np.pi / 10.0
    R = _euler_rotation_rad_f32(ax, ay, az).to(F64)                             # :453-454  [B,K,3,3]
    newU = torch.matmul(R, U * torch.ones_like(R))                              # :455
    sigmas = torch.matmul(torch.matmul(newU, torch.matmul(Ssc, Ssc)), newU.transpose(-1, -2))  # :456
    theta = theta_raw.to(F32) * 180.0                                           # :458
    return mus, sigmas, theta


# ---------------------------------------------------------------- synthetic data
def synth_inputs(B: int, K: int, seed: int = 56, angle: float = 0.0):
    """Synthetic renderer inputs drawn from the reference's activation ranges.

    mus = tanh(N(0,1)) (:378); Sigma from sigma_from_frame on tanh(N(0,1)) frames and
    N(0,1) pre-sigmoid scales (:383-401); roty = 180*tanh(N(0,1)) degrees (:458)
    [+ U(0, angle), :673]; rotx = rotz = 0 (:706).  seed 56 is the reference's own (:21).
    Returns dict of CPU tensors: mus [B,K,1,3] f32, sigma [B,K,3,3] f64, rotx/roty/rotz [B,1] f32.
    """
    g = torch.Generator(device="cpu").manual_seed(seed)
    mus = torch.tanh(torch.randn(B, K, 1, 3, generator=g, dtype=F32))
    v0 = torch.tanh(torch.randn(B, K, 3, generator=g, dtype=F32))
    v1 = torch.tanh(torch.randn(B, K, 3, generator=g, dtype=F32))
    u = torch.randn(B, K, 3, generator=g, dtype=F32)
    sigma = sigma_from_frame(v0, v1, u)
    roty = 180.0 * torch.tanh(torch.randn(B, 1, generator=g, dtype=F32))
    if angle:
        roty = roty + angle * torch.rand(B, 1, generator=g, dtype=F32)
    zr = torch.zeros_like(roty)
    return dict(mus=mus, sigma=sigma, rotx=zr.clone(), roty=roty, rotz=zr.clone())


def synth_target_masks(B: int, n: int, seed: int = 57) -> torch.Tensor:
    """Random binary discs in {-1,+1}, [B,n,n,1] f32 -- stand-in for the dataset masks (:672)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    c = torch.rand(B, 2, generator=g) * 1.0 - 0.5
    r = torch.rand(B, 1, generator=g) * 0.4 + 0.2
    ax = torch.from_numpy(tf_linspace_f32(n))
    yy, xx = torch.meshgrid(ax, ax, indexing="ij")
    d2 = (xx[None] - c[:, 0, None, None]) ** 2 + (yy[None] - c[:, 1, None, None]) ** 2
    m = (d2 < (r[:, :, None] ** 2)).to(F32) * 2.0 - 1.0
    return m.unsqueeze(-1)


def render_loss_and_grads(inp: dict, target_pm1: torch.Tensor, size: int = 256):
    """Forward + backward of the mask path exactly as the training graph wires it.

    get_landmarks -> sum over K at ``size`` -> mean|0.5*(mask+1) - m|
    (mask/mask_gen_dynamic.py:707, 785-787), gradients by autograd into mus, sigma, roty.
    Returns (mask [B,size,size] f32, loss, dmus f32, dsigma f64 (raw, NOT symmetrised), droty f32).
    """
    mus = inp["mus"].clone().requires_grad_(True)
    sigma = inp["sigma"].clone().requires_grad_(True)
    roty = inp["roty"].clone().requires_grad_(True)
    maps = get_landmarks(mus, sigma, inp["rotx"], roty, inp["rotz"], 0.0, 0.0, -2.0, 1.0, sizes=(size,))[0]
    loss = mask_recon_loss_bis(target_pm1, maps)
    loss.backward()
    return maps.detach().sum(-1), loss.detach(), mus.grad, sigma.grad, roty.grad


def sym(a: torch.Tensor) -> torch.Tensor:
    """Symmetric part over the last two axes (only sym(dL/dSigma) is observable upstream)."""
    return 0.5 * (a + a.transpose(-1, -2))


__all__ = [n for n in dir() if not n.startswith("_")]
