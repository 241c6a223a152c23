"""Direct consumers of the rendered maps, fused with the renderer.

* ``mask_recon_loss_bis`` -- ``mean|0.5*(mask+1) - sum_k g_k|``      mask/mask_gen_dynamic.py:785-787
* ``gm_cycle_loss``       -- ``mean|g_pose - g_cyc|`` over [B,n,n,K]  mask/mask_gen_dynamic.py:791
* ``colorize_landmark_maps`` / ``colorize_gaussians`` -- ``max_k g_k * COLORS[k]``   :221-231, infer_masks.py:174-184
* ``turntable``           -- the 60-angle inference loop of mask/infer_masks.py:426-463 as ONE batched launch

The two losses are differentiable scalars (``torch.autograd.Function``); neither ever writes a K-channel map.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from ._lib import GG_MOMENTS, GG_PARAM_STRIDE, LAYOUT_NHWC, check, int_array, lib, ptr_array
from .renderer import _prep_inputs, _ptr, _require_cuda, _stream, project

RGB_F32, RGB_U8, RGB_U8_INV = 0, 1, 2


def _project_params(mus3, sig, rots, t, focal, B, K, want_proj=False):
    dev = mus3.device
    params = torch.empty((B, K, GG_PARAM_STRIDE), dtype=torch.float32, device=dev)
    mu2d = sigma2d = None
    if want_proj:
        mu2d = torch.empty((B, K, 2), dtype=torch.float64, device=dev)
        sigma2d = torch.empty((B, K, 2, 2), dtype=torch.float64, device=dev)
    check(lib.gg_project_fwd(mus3.data_ptr(), sig.data_ptr(), int(sig.dtype == torch.float64),
                             rots[0].data_ptr(), rots[1].data_ptr(), rots[2].data_ptr(), *t, focal, B, K,
                             _ptr(mu2d), _ptr(sigma2d), params.data_ptr(), dev.index, _stream(dev)), "gg_project_fwd")
    return params, mu2d, sigma2d


def _project_adjoint(mus3, sig, rots, t, focal, B, K, partials, T, need):
    dev = mus3.device
    dmus = torch.empty((B, K, 3), dtype=torch.float32, device=dev) if need[0] else None
    dsig = torch.empty((B, K, 3, 3), dtype=torch.float64, device=dev) if need[1] else None
    drot = torch.empty((B, 3), dtype=torch.float32, device=dev) if any(need[2:5]) else None
    check(lib.gg_project_bwd(mus3.data_ptr(), sig.data_ptr(), int(sig.dtype == torch.float64),
                             rots[0].data_ptr(), rots[1].data_ptr(), rots[2].data_ptr(), *t, focal, B, K,
                             partials.data_ptr(), T, None, None, _ptr(dmus), _ptr(dsig), _ptr(drot),
                             dev.index, _stream(dev)), "gg_project_bwd")
    return dmus, dsig, drot


def _scale_grads(g, dmus, dsig, drot, sig_dtype, need):
    """Chain rule through the scalar loss: the kernels produced d loss / d input for grad_output = 1."""
    out = [None] * 5
    if need[0]:
        out[0] = dmus * g.to(dmus.dtype)
    if need[1]:
        out[1] = (dsig * g.to(dsig.dtype)).to(sig_dtype)
    for i in range(3):
        if need[2 + i]:
            out[2 + i] = (drot[:, i] * g.to(drot.dtype)).contiguous()
    return out


class _MaskReconLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mus3, sig, rx, ry, rz, target, cfg):
        t, focal, n = cfg
        B, K = mus3.shape[0], mus3.shape[1]
        dev = mus3.device
        rots = (rx, ry, rz)
        params, _, _ = _project_params(mus3, sig, rots, t, focal, B, K)
        sizes = int_array([n])
        T = check(lib.gg_splat_bwd_tiles(K, 1, sizes, int_array([0]), int_array([1]), LAYOUT_NHWC), "tiles")
        partials = torch.empty((B, T, K, GG_MOMENTS), dtype=torch.float32, device=dev)
        kind = 0 if target.dtype == torch.uint8 else 1
        if K <= 64:
            lp = torch.empty(B * T, dtype=torch.float32, device=dev)
            check(lib.gg_silhouette_loss_fused(params.data_ptr(), target.data_ptr(), kind, B, K, n, None,
                                               partials.data_ptr(), T, lp.data_ptr(), dev.index, _stream(dev)),
                  "gg_silhouette_loss_fused")
        else:                                   # K > 64: three launches, same arithmetic
            mask = torch.empty((B, n, n), dtype=torch.float32, device=dev)
            gm = torch.empty_like(mask)
            lp = torch.empty(296, dtype=torch.float32, device=dev)
            check(lib.gg_splat_fwd(params.data_ptr(), B, K, 1, sizes, ptr_array([None]), ptr_array([mask.data_ptr()]),
                                   LAYOUT_NHWC, dev.index, _stream(dev)), "gg_splat_fwd")
            check(lib.gg_mask_l1_loss_bwd(mask.data_ptr(), target.data_ptr(), kind, B, n, gm.data_ptr(), lp.data_ptr(),
                                          dev.index, _stream(dev)), "gg_mask_l1_loss_bwd")
            check(lib.gg_splat_bwd(params.data_ptr(), B, K, 1, sizes, ptr_array([None]), ptr_array([gm.data_ptr()]),
                                   LAYOUT_NHWC, partials.data_ptr(), T, dev.index, _stream(dev)), "gg_splat_bwd")
        ctx.save_for_backward(mus3, sig, rx, ry, rz, partials)
        ctx.cfg, ctx.T = cfg, T
        return (lp.double().sum() / float(B * n * n)).float()

    @staticmethod
    def backward(ctx, g):
        mus3, sig, rx, ry, rz, partials = ctx.saved_tensors
        t, focal, n = ctx.cfg
        B, K = mus3.shape[0], mus3.shape[1]
        need = ctx.needs_input_grad
        dmus, dsig, drot = _project_adjoint(mus3, sig, (rx, ry, rz), t, focal, B, K, partials, ctx.T, need)
        return (*_scale_grads(g, dmus, dsig, drot, sig.dtype, need), None, None)


def mask_recon_loss_bis(target, mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1., *, size: int = 256,
                        K: Optional[int] = None) -> torch.Tensor:
    """``reduce_mean(|(mask+1)*0.5 - reduce_sum(get_landmarks(...)[-1], -1)|)`` (mask/mask_gen_dynamic.py:785-787)
    as one differentiable scalar.  ``target`` is ``[B,size,size(,1)]``: uint8 = the raw 0..255 mask (normalised as
    ``mask/127.5 - 1``, :672) or float32 = the mask already in [-1, 1].  The silhouette is never written."""
    dev, B, Kk, mus3, sig, rots = _prep_inputs(mus, sigma, rotx, roty, rotz, K)
    _require_cuda(target)
    if target.numel() != B * size * size:
        raise ValueError("target must be [B,size,size] or [B,size,size,1]")
    if target.dtype not in (torch.uint8, torch.float32):
        target = target.float()
    target = target.reshape(B, size, size).contiguous()
    cfg = ((float(tx), float(ty), float(tz)), float(focal), int(size))
    return _MaskReconLoss.apply(mus3, sig, rots[0], rots[1], rots[2], target, cfg)


class _CycleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ma, sa, rxa, rya, rza, mb, sb, rxb, ryb, rzb, cfg):
        t, focal, n = cfg
        B, K = ma.shape[0], ma.shape[1]
        dev = ma.device
        pa, _, _ = _project_params(ma, sa, (rxa, rya, rza), t, focal, B, K)
        pb, _, _ = _project_params(mb, sb, (rxb, ryb, rzb), t, focal, B, K)
        T = check(lib.gg_cycle_tiles(n), "gg_cycle_tiles")
        nl = check(lib.gg_cycle_loss_count(B, K, n), "gg_cycle_loss_count")
        parta = torch.empty((B, T, K, GG_MOMENTS), dtype=torch.float32, device=dev)
        partb = torch.empty_like(parta)
        lp = torch.empty(nl, dtype=torch.float32, device=dev)
        check(lib.gg_cycle_l1_fused(pa.data_ptr(), pb.data_ptr(), B, K, n, parta.data_ptr(), partb.data_ptr(), T,
                                    lp.data_ptr(), dev.index, _stream(dev)), "gg_cycle_l1_fused")
        ctx.save_for_backward(ma, sa, rxa, rya, rza, mb, sb, rxb, ryb, rzb, parta, partb)
        ctx.cfg, ctx.T = cfg, T
        return (lp.double().sum() / float(B * n * n * K)).float()

    @staticmethod
    def backward(ctx, g):
        ma, sa, rxa, rya, rza, mb, sb, rxb, ryb, rzb, parta, partb = ctx.saved_tensors
        t, focal, n = ctx.cfg
        B, K = ma.shape[0], ma.shape[1]
        need = ctx.needs_input_grad
        ga = gb = [None] * 5
        if any(need[0:5]):
            ga = _scale_grads(g, *_project_adjoint(ma, sa, (rxa, rya, rza), t, focal, B, K, parta, ctx.T, need[0:5]),
                              sa.dtype, need[0:5])
        if any(need[5:10]):
            gb = _scale_grads(g, *_project_adjoint(mb, sb, (rxb, ryb, rzb), t, focal, B, K, partb, ctx.T, need[5:10]),
                              sb.dtype, need[5:10])
        return (*ga, *gb, None)


def gm_cycle_loss(render_a: Sequence[torch.Tensor], render_b: Sequence[torch.Tensor], tx, ty, tz, focal=1., *,
                  size: int = 256) -> torch.Tensor:
    """``reduce_mean(|get_landmarks(*a)[-1] - get_landmarks(*b)[-1]|)`` (mask/mask_gen_dynamic.py:791) as one
    differentiable scalar.  ``render_a`` / ``render_b`` = ``(mus, sigma, rotx, roty, rotz)`` of the two renders
    (same B and K).  Neither ``[B,size,size,K]`` tensor is materialised."""
    _, Ba, Ka, ma, sa, ra = _prep_inputs(*render_a, None)
    _, Bb, Kb, mb, sb, rb = _prep_inputs(*render_b, None)
    if (Ba, Ka) != (Bb, Kb):
        raise ValueError("the two renders must share B and K")
    cfg = ((float(tx), float(ty), float(tz)), float(focal), int(size))
    return _CycleLoss.apply(ma, sa, ra[0], ra[1], ra[2], mb, sb, rb[0], rb[1], rb[2], cfg)


# ---------------------------------------------------------------------------------------- viz composite
def _rgb_out(B, n, out, dev):
    if out == "f32":
        return torch.empty((B, n, n, 3), dtype=torch.float32, device=dev), RGB_F32
    if out == "u8":
        return torch.empty((B, n, n, 3), dtype=torch.uint8, device=dev), RGB_U8
    if out == "u8_inv":
        return torch.empty((B, n, n, 3), dtype=torch.uint8, device=dev), RGB_U8_INV
    raise ValueError("out must be 'f32', 'u8' or 'u8_inv'")


def colorize_landmark_maps(maps: torch.Tensor, colors: torch.Tensor, *, out: str = "f32") -> torch.Tensor:
    """Drop-in for ``colorize_landmark_maps`` (mask/mask_gen_dynamic.py:221-231; numpy twin
    mask/infer_masks.py:174-184): maps ``[B,H,W,K]``, colors ``[K,3]`` -> ``[B,H,W,3] = max_k maps[...,k]*colors[k]``.
    ``colors`` replaces the module global ``COLORS``.  ``out='u8_inv'`` gives ``255 - uint8(rgb*255)``
    (the frame written at infer_masks.py:463).  Viz only: not differentiable, like its use in the reference."""
    _require_cuda(maps, colors)
    B, H, W, K = maps.shape
    if H != W:
        raise ValueError("square maps only")
    maps = maps.detach().float().contiguous()
    colors = colors.detach().float().reshape(K, 3).contiguous()
    o, mode = _rgb_out(B, H, out, maps.device)
    check(lib.gg_colorize(None, maps.data_ptr(), colors.data_ptr(), B, K, H, o.data_ptr(), mode, maps.device.index,
                          _stream(maps.device)), "gg_colorize")
    return o


def colorize_gaussians(mus, sigma, rotx, roty, rotz, tx, ty, tz, colors, focal=1., *, size: int = 256,
                       out: str = "f32"):
    """``colorize_landmark_maps(get_landmarks(...)[-1])`` without writing the maps."""
    dev, B, K, mus3, sig, rots = _prep_inputs(mus, sigma, rotx, roty, rotz, None)
    params, _, _ = _project_params(mus3.detach(), sig.detach(), [r.detach() for r in rots],
                                   (float(tx), float(ty), float(tz)), float(focal), B, K)
    colors = colors.detach().float().reshape(K, 3).contiguous().to(dev)
    o, mode = _rgb_out(B, size, out, dev)
    check(lib.gg_colorize(params.data_ptr(), None, colors.data_ptr(), B, K, size, o.data_ptr(), mode, dev.index,
                          _stream(dev)), "gg_colorize")
    return o


def turntable(mu3d: torch.Tensor, sig3d: torch.Tensor, theta0, angles, colors: Optional[torch.Tensor] = None, *,
              tx=0., ty=0., tz=-2., focal=1., size: int = 256, want_maps: bool = False):
    """The inference loop of mask/infer_masks.py:426-463 in one batched launch per kernel.

    ``mu3d [1,K,1,3]`` f32 and ``sig3d [1,K,3,3]`` f64 are the tensors the ``z_to_l`` graph emits (and that the
    reference saves as ``m_%d.npy`` / ``sig3d_%d.npy``, :447-448); ``theta0`` the predicted yaw (degrees),
    ``angles`` the F turntable offsets (reference: ``np.linspace(0, angle, 60)``, :428).  Returns a dict with
    ``mu2d [F,K,2]``, ``sigma2d [F,K,2,2]`` (float32, what is fed to the frozen ``l_to_m`` decoder, :455-456),
    optionally ``maps [F,size,size,K]`` and, when ``colors`` is given, ``frames [F,size,size,3]`` uint8
    ``= 255 - uint8(colorize*255)`` (:463).  The reference opens a new TF session per angle; here all F angles are
    one batch."""
    _require_cuda(mu3d, sig3d)
    dev = mu3d.device
    ang = torch.as_tensor(angles, dtype=torch.float32, device=dev).reshape(-1)
    F = ang.numel()
    K = sig3d.shape[-3]
    roty = (torch.as_tensor(theta0, dtype=torch.float32, device=dev).reshape(-1)[:1] + ang).reshape(F, 1)   # thet3d[0,0] + x
    zr = torch.zeros_like(roty)
    mus = mu3d.reshape(1, K, 1, 3).expand(F, K, 1, 3).contiguous()
    sig = sig3d.reshape(1, K, 3, 3).expand(F, K, 3, 3).contiguous()
    mu2d, sigma2d = project(mus, sig, zr, roty, zr, tx, ty, tz, focal)
    res = {"mu2d": mu2d.float(), "sigma2d": sigma2d.float(), "roty": roty}
    if want_maps:
        from .renderer import get_gaussian_maps_2d
        res["maps"] = get_gaussian_maps_2d(res["mu2d"], res["sigma2d"], [size, size])
    if colors is not None:
        # the reference colourises the float32-raster maps of the inference flavour (infer_masks.py:265, 463)
        params = torch.empty((F, K, GG_PARAM_STRIDE), dtype=torch.float32, device=dev)
        check(lib.gg_conic_fwd(res["mu2d"].data_ptr(), res["sigma2d"].data_ptr(), 0, F, K, params.data_ptr(),
                               dev.index, _stream(dev)), "gg_conic_fwd")
        cols = colors.detach().float().reshape(K, 3).contiguous().to(dev)
        frames = torch.empty((F, size, size, 3), dtype=torch.uint8, device=dev)
        check(lib.gg_colorize(params.data_ptr(), None, cols.data_ptr(), F, K, size, frames.data_ptr(), RGB_U8_INV,
                              dev.index, _stream(dev)), "gg_colorize")
        res["frames"] = frames
    return res
