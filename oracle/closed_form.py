"""TEST INFRASTRUCTURE ONLY -- independent numpy float64 closed form of the renderer.

A second oracle, written from the geometry rather than from the reference's op
list, used to cross-check ``reference_port.py`` (and, through it, the CUDA
kernels, which implement this same closed form in float64 registers):

* forward: dual-conic projection of the 1-sigma ellipsoid,
  algebraically equal to mask/mask_gen_dynamic.py:274-309 for symmetric Sigma;
* backward: hand-derived chain rule from per-Gaussian pixel
  moments to d mus, d Sigma (symmetric part), d angles.

Not imported by ``gaussigan_b200``; see ``oracle/__init__.py``.
"""
from __future__ import annotations

import numpy as np

from .reference_port import tf_linspace_f32


def euler_R(rx_deg, ry_deg, rz_deg):
    """R = RX @ RY @ RZ with the reference's column-stack convention
    (mask/mask_gen_dynamic.py:240-255).  Angles [B] in degrees; trig in float32
    like the reference, result float64 [B,3,3].  Also returns dR/d(angle_rad)."""
    def rad(a):
        return (np.asarray(a, np.float32) * np.float32(np.pi) / np.float32(180.0)).astype(np.float32)
    ax, ay, az = rad(rx_deg), rad(ry_deg), rad(rz_deg)
    z = np.zeros_like(ax, dtype=np.float64)
    o = np.ones_like(z)

    def mat(rows):
        return np.stack([np.stack(r, -1) for r in rows], -2)

    def cs(a):   # float32 trig, correctly rounded (reference_port.cos_f32)
        a64 = a.astype(np.float64)
        return (np.cos(a64).astype(np.float32).astype(np.float64),
                np.sin(a64).astype(np.float32).astype(np.float64))
    cx, sx = cs(ax); cy, sy = cs(ay); cz, sz = cs(az)
    RX = mat([[o, z, z], [z, cx, sx], [z, -sx, cx]]); dRX = mat([[z, z, z], [z, -sx, cx], [z, -cx, -sx]])
    RY = mat([[cy, z, -sy], [z, o, z], [sy, z, cy]]); dRY = mat([[-sy, z, -cy], [z, z, z], [cy, z, -sy]])
    RZ = mat([[cz, sz, z], [-sz, cz, z], [z, z, o]]); dRZ = mat([[-sz, cz, z], [-cz, -sz, z], [z, z, z]])
    # the reference multiplies in float32; round the product the same way
    def mm32(a, b):   # pinned order, products and sums rounded to float32 (reference_port.matmul3_f32)
        p = (a[..., :, :, None] * b[..., None, :, :]).astype(np.float32)
        return ((p[..., 0, :] + p[..., 1, :]).astype(np.float32) + p[..., 2, :]).astype(np.float32)
    R32 = mm32(mm32(RX.astype(np.float32), RY.astype(np.float32)), RZ.astype(np.float32))
    R = R32.astype(np.float64)
    return R, (dRX @ RY @ RZ, RX @ dRY @ RZ, RX @ RY @ dRZ)


def project_closed(mus, sigma, rx, ry, rz, t=(0.0, 0.0, -2.0), focal=1.0):
    """mus [B,K,3], sigma [B,K,3,3], angles [B] deg -> (mu2d [B,K,2], sigma2d [B,K,2,2], cache)."""
    mus = np.asarray(mus, np.float64)
    S = np.asarray(sigma, np.float64)
    S = 0.5 * (S + np.swapaxes(S, -1, -2))
    R, dR = euler_R(rx, ry, rz)
    Rb = R[:, None]                                               # [B,1,3,3]
    Sp = Rb @ S @ np.swapaxes(Rb, -1, -2)
    m = np.einsum("bij,bkj->bki", R, mus) + np.asarray(t, np.float64)
    mxy, mz = m[..., :2], m[..., 2]
    w = mz * mz - Sp[..., 2, 2]
    beta = Sp[..., :2, 2] - mxy * mz[..., None]
    c0 = beta / w[..., None]
    f = float(np.float32(focal))
    mu2d = f * c0
    A = mxy[..., :, None] * mxy[..., None, :] - Sp[..., :2, :2]
    sigma2d = c0[..., :, None] * c0[..., None, :] - A / w[..., None, None]
    cache = dict(R=R, dR=dR, S=S, mus=mus, m=m, w=w, beta=beta, c0=c0, A=A, f=f)
    return mu2d, sigma2d, cache


def conic(sigma2d):
    """P = inv(sigma2d), returned as (a, b, c) with d = a dx^2 + 2 b dx dy + c dy^2."""
    s00, s01, s10, s11 = sigma2d[..., 0, 0], sigma2d[..., 0, 1], sigma2d[..., 1, 0], sigma2d[..., 1, 1]
    det = s00 * s11 - s01 * s10
    return s11 / det, -0.5 * (s01 + s10) / det, s00 / det


def raster(mu2d, sigma2d, n):
    """[B,K,2],[B,K,2,2] -> maps [B,n,n,K] float64 on the reference grid (:121-141)."""
    ax = tf_linspace_f32(n).astype(np.float64)
    a, b, c = conic(sigma2d)
    dx = ax[None, None, None, :] - mu2d[..., 0][..., None, None]      # [B,K,1,n]
    dy = ax[None, None, :, None] - mu2d[..., 1][..., None, None]      # [B,K,n,1]
    d = (a[..., None, None] * dx * dx + 2.0 * b[..., None, None] * dx * dy + c[..., None, None] * dy * dy)
    return np.transpose(np.exp(-d), (0, 2, 3, 1))


def pixel_moments(mu2d, sigma2d, gbar_maps):
    """Sum over scales of the five per-Gaussian moments of w = g * Gbar.

    gbar_maps: list of [B,n,n,K] upstream gradients (one per scale).
    Returns [B,K,5] = (M20, M11, M02, M10, M01) in (dx, dy) coordinates.
    """
    a, b, c = conic(sigma2d)
    out = np.zeros(mu2d.shape[:2] + (5,), np.float64)
    for gb in gbar_maps:
        n = gb.shape[1]
        ax = tf_linspace_f32(n).astype(np.float64)
        dx = ax[None, None, None, :] - mu2d[..., 0][..., None, None]
        dy = ax[None, None, :, None] - mu2d[..., 1][..., None, None]
        d = (a[..., None, None] * dx * dx + 2.0 * b[..., None, None] * dx * dy + c[..., None, None] * dy * dy)
        wgt = np.exp(-d) * np.transpose(np.asarray(gb, np.float64), (0, 3, 1, 2))
        out[..., 0] += (wgt * dx * dx).sum((-1, -2))
        out[..., 1] += (wgt * dx * dy).sum((-1, -2))
        out[..., 2] += (wgt * dy * dy).sum((-1, -2))
        out[..., 3] += (wgt * dx * np.ones_like(dy)).sum((-1, -2))
        out[..., 4] += (wgt * dy * np.ones_like(dx)).sum((-1, -2))
    return out


def conic_backward(sigma2d, mom):
    """Moments -> (d mu2d [B,K,2], d sigma2d [B,K,2,2] symmetric)."""
    a, b, c = conic(sigma2d)
    P = np.stack([np.stack([a, b], -1), np.stack([b, c], -1)], -2)
    Pbar = -np.stack([np.stack([mom[..., 0], mom[..., 1]], -1),
                      np.stack([mom[..., 1], mom[..., 2]], -1)], -2)
    dmu = 2.0 * np.einsum("...ij,...j->...i", P, mom[..., 3:5])
    dsig = -P @ Pbar @ P
    return dmu, dsig


def project_backward(cache, dmu2d, dsig2d):
    """(d mu2d, d sigma2d) -> d mus [B,K,3], d Sigma [B,K,3,3] (symmetric), d angles_deg 3 x [B]."""
    R, S, mus, m, w, beta, c0, A, f = (cache[k] for k in ("R", "S", "mus", "m", "w", "beta", "c0", "A", "f"))
    mxy, mz = m[..., :2], m[..., 2]
    c0bar = 2.0 * np.einsum("...ij,...j->...i", dsig2d, c0) + f * dmu2d
    Abar = -dsig2d / w[..., None, None]
    wbar = ((dsig2d * A).sum((-1, -2)) - (c0bar * beta).sum(-1)) / (w * w)
    bbar = c0bar / w[..., None]
    mbar = np.zeros_like(m)
    mbar[..., :2] = 2.0 * np.einsum("...ij,...j->...i", Abar, mxy) - bbar * mz[..., None]
    mbar[..., 2] = -(bbar * mxy).sum(-1) + 2.0 * mz * wbar
    Sbar = np.zeros_like(S)
    Sbar[..., :2, :2] = -Abar
    Sbar[..., :2, 2] = 0.5 * bbar
    Sbar[..., 2, :2] = 0.5 * bbar
    Sbar[..., 2, 2] = -wbar
    Rb = R[:, None]
    dmus = np.einsum("bji,bkj->bki", R, mbar)
    dSigma = np.swapaxes(Rb, -1, -2) @ Sbar @ Rb
    Rbar = (mbar[..., :, None] * mus[..., None, :] + 2.0 * Sbar @ Rb @ S).sum(1)   # [B,3,3]
    dang = [np.float64(np.pi / 180.0) * (Rbar * d).sum((-1, -2)) for d in cache["dR"]]
    return dmus, dSigma, dang


def sphere_kat(s2: float = 0.09, tz: float = -2.0):
    """Known answer: sphere Sigma = s2*I on the optical axis at distance |tz| -> mu2d = 0,
    sigma2d = s2 / (tz^2 - s2) * I  (SURVEY.md section 4)."""
    return np.zeros(2), (s2 / (tz * tz - s2)) * np.eye(2)
