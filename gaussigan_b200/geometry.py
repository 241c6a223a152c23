"""Gaussian construction and per-frame deformation: the two small geometry stages in front of the renderer.

* ``sigma_from_frame(v0, v1, u)``  -- tail of ``get_musigma`` (mask/mask_gen_dynamic.py:383-401)
* ``deform(mu_t, scale, rx, ry, rz, theta, cannmu, cannsigma)`` -- tail of the dynamic ``transform_mu_sigma``
  (mask/mask_gen_dynamic.py:426-458); the SVD of the canonical covariances (:430) is a Jacobi kernel
  (``svd3``: no cuSOLVER, no library fallback), everything after it is one CUDA launch per direction.

Both take the OUTPUTS of the dense heads (the MLPs are out of scope) and are differentiable.
"""
from __future__ import annotations

import torch

from ._lib import check, lib
from .renderer import _ptr, _require_cuda, _stream


class _SigmaFromFrame(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v0, v1, u):
        N = v0.numel() // 3
        sigma = torch.empty(v0.shape[:-1] + (3, 3), dtype=torch.float64, device=v0.device)
        check(lib.gg_sigma_from_frame_fwd(v0.data_ptr(), v1.data_ptr(), u.data_ptr(), N, sigma.data_ptr(),
                                          v0.device.index, _stream(v0.device)), "gg_sigma_from_frame_fwd")
        ctx.save_for_backward(v0, v1, u)
        return sigma

    @staticmethod
    def backward(ctx, g):
        v0, v1, u = ctx.saved_tensors
        N = v0.numel() // 3
        g = g.to(torch.float64).contiguous()
        dv0, dv1, du = torch.empty_like(v0), torch.empty_like(v1), torch.empty_like(u)
        check(lib.gg_sigma_from_frame_bwd(v0.data_ptr(), v1.data_ptr(), u.data_ptr(), g.data_ptr(), N, dv0.data_ptr(),
                                          dv1.data_ptr(), du.data_ptr(), v0.device.index, _stream(v0.device)),
              "gg_sigma_from_frame_bwd")
        return dv0, dv1, du


def sigma_from_frame(v0_raw: torch.Tensor, v1_raw: torch.Tensor, u_raw: torch.Tensor) -> torch.Tensor:
    """``v0_raw, v1_raw`` = tanh head outputs ``[N,K,3]``, ``u_raw`` = pre-sigmoid head output ``[N,K,3]`` (float32)
    -> ``Sigma [N,K,3,3]`` float64 (built in float32 like the reference, :400-401)."""
    _require_cuda(v0_raw, v1_raw, u_raw)
    if not (v0_raw.shape == v1_raw.shape == u_raw.shape) or v0_raw.shape[-1] != 3:
        raise ValueError("v0_raw, v1_raw, u_raw must share a [...,3] shape")
    f = lambda t: t.float().contiguous()
    return _SigmaFromFrame.apply(f(v0_raw), f(v1_raw), f(u_raw))


class _SVD3(torch.autograd.Function):
    """``S, U, _ = tf.linalg.svd(A)`` for 3x3 matrices (mask/mask_gen_dynamic.py:430) as a CUDA kernel: no host
    synchronisation (``torch.linalg.svd`` checks cuSOLVER's status on the host), so it can live in a CUDA graph.
    Returns ``(U, S)``; like the reference, V carries no gradient."""

    @staticmethod
    def forward(ctx, A):
        N = A.numel() // 9
        U, V = torch.empty_like(A), torch.empty_like(A)
        S = torch.empty(A.shape[:-2] + (3,), dtype=torch.float64, device=A.device)
        check(lib.gg_svd3_fwd(A.data_ptr(), N, U.data_ptr(), S.data_ptr(), V.data_ptr(), A.device.index,
                              _stream(A.device)), "gg_svd3_fwd")
        ctx.save_for_backward(U, S, V)
        ctx.set_materialize_grads(False)
        return U, S

    @staticmethod
    def backward(ctx, gU, gS):
        U, S, V = ctx.saved_tensors
        N = U.numel() // 9
        gU = None if gU is None else gU.to(torch.float64).contiguous()
        gS = None if gS is None else gS.to(torch.float64).contiguous()
        gA = torch.empty_like(U)
        check(lib.gg_svd3_bwd(U.data_ptr(), S.data_ptr(), V.data_ptr(), _ptr(gU), _ptr(gS), N, gA.data_ptr(),
                              U.device.index, _stream(U.device)), "gg_svd3_bwd")
        return gA


def svd3(A: torch.Tensor):
    """``A [...,3,3]`` float64, CUDA -> ``(U [...,3,3], S [...,3])``, the left singular vectors and the singular values
    (descending) of each matrix; differentiable (square full-rank formula, no gradient through V); never
    synchronises with the host."""
    _require_cuda(A)
    if A.shape[-2:] != (3, 3):
        raise ValueError("svd3 expects [...,3,3]")
    return _SVD3.apply(A.to(torch.float64).contiguous())


class _Deform(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mu_t, scale, ang, theta_raw, cmu, U, S):
        B, K = mu_t.shape[0], mu_t.shape[1]
        dev = mu_t.device
        mus = torch.empty((B, K, 3), dtype=torch.float32, device=dev)
        sig = torch.empty((B, K, 3, 3), dtype=torch.float64, device=dev)
        theta = torch.empty((B,), dtype=torch.float32, device=dev)
        check(lib.gg_deform_fwd(mu_t.data_ptr(), scale.data_ptr(), ang.data_ptr(), cmu.data_ptr(), U.data_ptr(),
                                S.data_ptr(), theta_raw.data_ptr(), B, K, mus.data_ptr(), sig.data_ptr(),
                                theta.data_ptr(), dev.index, _stream(dev)), "gg_deform_fwd")
        ctx.save_for_backward(scale, ang, U, S)
        ctx.set_materialize_grads(False)
        return mus, sig, theta

    @staticmethod
    def backward(ctx, gmus, gsig, gtheta):
        scale, ang, U, S = ctx.saved_tensors
        B, K = scale.shape[0], scale.shape[1]
        dev = scale.device
        gmus = None if gmus is None else gmus.float().contiguous()
        gsig = None if gsig is None else gsig.to(torch.float64).contiguous()
        gtheta = None if gtheta is None else gtheta.float().contiguous()
        f32 = dict(dtype=torch.float32, device=dev)
        dth, dmu_t, dsc, dang, dcmu = (torch.empty(s, **f32) for s in ((B,), (B, K, 3), (B, K, 3), (B, K, 3), (B, K, 3)))
        dU = torch.empty((B, K, 3, 3), dtype=torch.float64, device=dev)
        dS = torch.empty((B, K, 3), dtype=torch.float64, device=dev)
        check(lib.gg_deform_bwd(scale.data_ptr(), ang.data_ptr(), U.data_ptr(), S.data_ptr(), _ptr(gmus), _ptr(gsig),
                                _ptr(gtheta), B, K, dth.data_ptr(), dmu_t.data_ptr(), dsc.data_ptr(), dang.data_ptr(),
                                dcmu.data_ptr(), dU.data_ptr(), dS.data_ptr(), dev.index, _stream(dev)), "gg_deform_bwd")
        # canonical parameters are broadcast over the batch (:428, :455): their gradient is the batch sum
        return dmu_t, dsc, dang, dth, dcmu.sum(0), dU.sum(0), dS.sum(0)


def deform(mu_t_raw, scale_raw, rx_raw, ry_raw, rz_raw, theta_raw, cannmu, cannsigma, factors=None):
    """mask/mask_gen_dynamic.py:426-458.  Head outputs AFTER tanh: ``mu_t_raw [B,K*3]``, ``scale_raw [B,K*3]``,
    ``r*_raw [B,K]``, ``theta_raw [B,1]``; canonical ``cannmu [1,K,1,3]`` f32, ``cannsigma [1,K,3,3]`` f64.
    Returns ``mus [B,K,1,3]`` f32, ``sigmas [B,K,3,3]`` f64, ``theta [B,1]`` f32 degrees -- ready for ``get_landmarks``.
    The canonical covariances are factorised (:430) by the library's own Jacobi kernel (``svd3``).  ``factors=(U, S)``
    lets a caller that already holds a factorisation of ``cannsigma`` (e.g. a test feeding the same U, S to both sides
    of a comparison) pass it in; Sigma_b does not depend on the column signs of U."""
    _require_cuda(mu_t_raw, scale_raw, rx_raw, ry_raw, rz_raw, theta_raw, cannmu, cannsigma)
    K = cannsigma.shape[-3]
    B = mu_t_raw.shape[0]
    if factors is None:      # :430, Jacobi kernel (graph-capturable)
        U, S = svd3(cannsigma.to(torch.float64).reshape(K, 3, 3))
    else:
        U, S = factors
        U, S = U.to(torch.float64).reshape(K, 3, 3), S.to(torch.float64).reshape(K, 3)
    ang = torch.stack([rx_raw.reshape(B, K), ry_raw.reshape(B, K), rz_raw.reshape(B, K)], dim=-1)
    f = lambda t: t.float().contiguous()
    mus, sig, theta = _Deform.apply(f(mu_t_raw.reshape(B, K, 3)), f(scale_raw.reshape(B, K, 3)), f(ang),
                                    f(theta_raw.reshape(B)), f(cannmu.reshape(K, 3)), U.contiguous(), S.contiguous())
    return mus.reshape(B, K, 1, 3), sig, theta.reshape(B, 1)
