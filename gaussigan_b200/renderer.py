"""Host-side mirror of the reference's renderer interface, backed by libgaussigan_sm100.so.

Same names, argument meaning and error behaviour as the reference's Python functions:

* ``get_landmarks(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1.)`` -- mask/mask_gen_dynamic.py:233-323
  (and the five other copies): returns the list of four ``[B,n,n,K]`` float32 maps.
* ``get_gaussian_maps_2d(mu, sigma, shape_hw)`` -- mask/mask_gen_dynamic.py:112-143.
* ``get_landmarks_infer(...)`` -- mask/infer_masks.py:186-267: ``(mu2d, sigma2d, maps256)`` in float32.
* ``render_silhouette(...)`` -- get_landmarks followed by the sum composite of
  mask/mask_gen_dynamic.py:785-786, computed without materialising the K-channel maps.

All are differentiable (``torch.autograd.Function`` over the C ABI).  Inputs must be CUDA tensors;
there is no CPU path.  PyTorch supplies device memory and streams only.
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import GG_MOMENTS, GG_PARAM_STRIDE, LAYOUT_NCHW, LAYOUT_NHWC, check, int_array, lib, ptr_array

#: the reference's hard-coded pyramid (mask/mask_gen_dynamic.py:236, 315-316)
REFERENCE_SIZES = (32, 64, 128, 256)


class _Spec(NamedTuple):
    tx: float
    ty: float
    tz: float
    focal: float
    sizes: Tuple[int, ...]
    want_maps: Tuple[bool, ...]
    want_masks: Tuple[bool, ...]
    layout: int
    want_proj: bool
    views: int = 1          # cameras per set of Gaussians (rot* hold views * B angles, view-major)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda(*ts: torch.Tensor) -> torch.device:
    dev = ts[0].device
    for t in ts:
        if not t.is_cuda:
            raise RuntimeError("gaussigan_b200 runs on CUDA (sm_100a) only: got a %s tensor; there is no CPU path"
                               % t.device.type)
        if t.device != dev:
            raise RuntimeError("all renderer inputs must live on the same device")
    return dev


def _layout_code(layout: str) -> int:
    if layout == "nhwc":
        return LAYOUT_NHWC
    if layout == "nchw":
        return LAYOUT_NCHW
    raise ValueError("layout must be 'nhwc' (reference) or 'nchw'")


def _splat_forward(params: torch.Tensor, B: int, K: int, spec: _Spec):
    dev = params.device
    maps: List[Optional[torch.Tensor]] = []
    masks: List[Optional[torch.Tensor]] = []
    for n, wm, wk in zip(spec.sizes, spec.want_maps, spec.want_masks):
        shape = (B, n, n, K) if spec.layout == LAYOUT_NHWC else (B, K, n, n)
        maps.append(torch.empty(shape, dtype=torch.float32, device=dev) if wm else None)
        masks.append(torch.empty((B, n, n), dtype=torch.float32, device=dev) if wk else None)
    check(lib.gg_splat_fwd(params.data_ptr(), B, K, len(spec.sizes), int_array(spec.sizes),
                           ptr_array([_ptr(m) for m in maps]), ptr_array([_ptr(m) for m in masks]),
                           spec.layout, dev.index, _stream(dev)), "gg_splat_fwd")
    return maps, masks


def _splat_backward(params: torch.Tensor, B: int, K: int, spec: _Spec,
                    gmaps: Sequence[Optional[torch.Tensor]], gmasks: Sequence[Optional[torch.Tensor]]):
    """-> (partials [B,T,K,5] or None, T)"""
    dev = params.device
    hm = [g is not None for g in gmaps]
    hk = [g is not None for g in gmasks]
    if not any(hm) and not any(hk):
        return None, 0
    sizes = int_array(spec.sizes)
    T = check(lib.gg_splat_bwd_tiles(K, len(spec.sizes), sizes, int_array(hm), int_array(hk), spec.layout),
              "gg_splat_bwd_tiles")
    partials = torch.empty((B, T, K, GG_MOMENTS), dtype=torch.float32, device=dev)
    check(lib.gg_splat_bwd(params.data_ptr(), B, K, len(spec.sizes), sizes,
                           ptr_array([_ptr(g) for g in gmaps]), ptr_array([_ptr(g) for g in gmasks]),
                           spec.layout, partials.data_ptr(), T, dev.index, _stream(dev)), "gg_splat_bwd")
    return partials, T


def _clean_grad(g: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if g is None:
        return None
    if g.dtype != torch.float32:
        g = g.float()
    return g.contiguous()


class _Render(torch.autograd.Function):
    """(mus [B,K,3] f32, sigma [B,K,3,3], rot* [B] f32 deg) -> [mu2d, sigma2d,] maps..., masks..."""

    @staticmethod
    def forward(ctx, mus, sigma, rotx, roty, rotz, spec: _Spec):
        dev = mus.device
        Bg, K, V = mus.shape[0], mus.shape[1], spec.views
        B = Bg * V                                   # rendered images: view-major, image v * Bg + b = set b, camera v
        params = torch.empty((B, K, GG_PARAM_STRIDE), dtype=torch.float32, device=dev)
        mu2d = sigma2d = None
        if spec.want_proj:
            mu2d = torch.empty((B, K, 2), dtype=torch.float64, device=dev)
            sigma2d = torch.empty((B, K, 2, 2), dtype=torch.float64, device=dev)
        check(lib.gg_project_fwd_views(mus.data_ptr(), sigma.data_ptr(), int(sigma.dtype == torch.float64),
                                       rotx.data_ptr(), roty.data_ptr(), rotz.data_ptr(),
                                       spec.tx, spec.ty, spec.tz, spec.focal, Bg, V, K,
                                       _ptr(mu2d), _ptr(sigma2d), params.data_ptr(), dev.index, _stream(dev)),
              "gg_project_fwd_views")
        maps, masks = ([], [])
        if any(spec.want_maps) or any(spec.want_masks):
            maps, masks = _splat_forward(params, B, K, spec)
        ctx.save_for_backward(mus, sigma, rotx, roty, rotz, params)
        ctx.spec = spec
        ctx.set_materialize_grads(False)
        outs = []
        if spec.want_proj:
            outs += [mu2d, sigma2d]
        outs += [m for m in maps if m is not None]
        outs += [m for m in masks if m is not None]
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        mus, sigma, rotx, roty, rotz, params = ctx.saved_tensors
        spec: _Spec = ctx.spec
        dev = mus.device
        Bg, K, V = mus.shape[0], mus.shape[1], spec.views
        B = Bg * V
        it = iter(grads)
        gmu2d = gsig2d = None
        if spec.want_proj:
            gmu2d, gsig2d = next(it), next(it)
            gmu2d = None if gmu2d is None else gmu2d.to(torch.float64).contiguous()
            gsig2d = None if gsig2d is None else gsig2d.to(torch.float64).contiguous()
        gmaps = [_clean_grad(next(it)) if w else None for w in spec.want_maps]
        gmasks = [_clean_grad(next(it)) if w else None for w in spec.want_masks]
        partials, T = _splat_backward(params, B, K, spec, gmaps, gmasks)
        if partials is None and gmu2d is None and gsig2d is None:
            return None, None, None, None, None, None
        need = ctx.needs_input_grad
        dmus = torch.empty((Bg, K, 3), dtype=torch.float32, device=dev) if need[0] else None
        dsigma = torch.empty((Bg, K, 3, 3), dtype=torch.float64, device=dev) if need[1] else None
        drot = torch.empty((B, 3), dtype=torch.float32, device=dev) if any(need[2:5]) else None
        check(lib.gg_project_bwd_views(mus.data_ptr(), sigma.data_ptr(), int(sigma.dtype == torch.float64),
                                       rotx.data_ptr(), roty.data_ptr(), rotz.data_ptr(),
                                       spec.tx, spec.ty, spec.tz, spec.focal, Bg, V, K,
                                       _ptr(partials), T, _ptr(gmu2d), _ptr(gsig2d),
                                       _ptr(dmus), _ptr(dsigma), _ptr(drot), dev.index, _stream(dev)),
              "gg_project_bwd_views")
        if dsigma is not None and sigma.dtype != torch.float64:
            dsigma = dsigma.to(sigma.dtype)
        return (dmus, dsigma,
                drot[:, 0].contiguous() if need[2] else None,
                drot[:, 1].contiguous() if need[3] else None,
                drot[:, 2].contiguous() if need[4] else None,
                None)


class _Raster(torch.autograd.Function):
    """(mu2d [B,K,2], sigma2d [B,K,2,2]) -> maps [B,n,n,K] / [B,K,n,n] f32"""

    @staticmethod
    def forward(ctx, mu, sigma, spec: _Spec):
        dev = mu.device
        B, K = mu.shape[0], mu.shape[1]
        params = torch.empty((B, K, GG_PARAM_STRIDE), dtype=torch.float32, device=dev)
        check(lib.gg_conic_fwd(mu.data_ptr(), sigma.data_ptr(), int(mu.dtype == torch.float64), B, K,
                               params.data_ptr(), dev.index, _stream(dev)), "gg_conic_fwd")
        maps, masks = _splat_forward(params, B, K, spec)
        ctx.save_for_backward(mu, sigma, params)
        ctx.spec = spec
        ctx.set_materialize_grads(False)
        return tuple([m for m in maps if m is not None] + [m for m in masks if m is not None])

    @staticmethod
    def backward(ctx, *grads):
        mu, sigma, params = ctx.saved_tensors
        spec: _Spec = ctx.spec
        dev = mu.device
        B, K = mu.shape[0], mu.shape[1]
        it = iter(grads)
        gmaps = [_clean_grad(next(it)) if w else None for w in spec.want_maps]
        gmasks = [_clean_grad(next(it)) if w else None for w in spec.want_masks]
        partials, T = _splat_backward(params, B, K, spec, gmaps, gmasks)
        if partials is None:
            return None, None, None
        dmu = torch.empty((B, K, 2), dtype=torch.float64, device=dev)
        dsig = torch.empty((B, K, 2, 2), dtype=torch.float64, device=dev)
        check(lib.gg_conic_bwd(mu.data_ptr(), sigma.data_ptr(), int(mu.dtype == torch.float64), B, K,
                               partials.data_ptr(), T, dmu.data_ptr(), dsig.data_ptr(), dev.index, _stream(dev)),
              "gg_conic_bwd")
        return dmu.to(mu.dtype), dsig.to(sigma.dtype), None


class _RenderInto(torch.autograd.Function):
    """(mus, sigma, rot*, spec, *bufs) -> bufs, with the K maps of scale i written IN PLACE into channels
    [c0_i, c0_i + K) of the NHWC buffer bufs[i] (``gg_splat_fwd_concat``); the backward reads the same channel slice of
    the buffers' gradients in place (``gg_splat_bwd_concat``) and hands the gradients on unchanged for the channels
    this op did not write."""

    @staticmethod
    def forward(ctx, mus, sigma, rotx, roty, rotz, spec, offsets, *bufs):
        dev = mus.device
        B, K = mus.shape[0], mus.shape[1]
        params = torch.empty((B, K, GG_PARAM_STRIDE), dtype=torch.float32, device=dev)
        check(lib.gg_project_fwd(mus.data_ptr(), sigma.data_ptr(), int(sigma.dtype == torch.float64),
                                 rotx.data_ptr(), roty.data_ptr(), rotz.data_ptr(),
                                 spec.tx, spec.ty, spec.tz, spec.focal, B, K,
                                 None, None, params.data_ptr(), dev.index, _stream(dev)), "gg_project_fwd")
        strides = int_array([b.shape[-1] for b in bufs])
        offs = int_array(offsets)
        check(lib.gg_splat_fwd_concat(params.data_ptr(), B, K, len(bufs), int_array(spec.sizes),
                                      ptr_array([b.data_ptr() for b in bufs]), strides, offs, None, dev.index,
                                      _stream(dev)), "gg_splat_fwd_concat")
        ctx.mark_dirty(*bufs)
        ctx.save_for_backward(mus, sigma, rotx, roty, rotz, params)
        ctx.spec, ctx.offsets, ctx.cstrides = spec, tuple(offsets), tuple(b.shape[-1] for b in bufs)
        ctx.set_materialize_grads(False)
        return bufs

    @staticmethod
    def backward(ctx, *gbufs):
        mus, sigma, rotx, roty, rotz, params = ctx.saved_tensors
        spec: _Spec = ctx.spec
        dev = mus.device
        B, K = mus.shape[0], mus.shape[1]
        gb = [None if g is None else _clean_grad(g) for g in gbufs]
        hm = [g is not None for g in gb]
        none = (None,) * 7
        if not any(hm):
            return none + tuple(gbufs)
        sizes, strides, offs = int_array(spec.sizes), int_array(ctx.cstrides), int_array(ctx.offsets)
        # In place (128-bit loads of the channel slice) when the slice is a run of aligned quads of Gaussians, K = 4 * 2^m:
        # 2.5x faster than slicing first (measured, tools/time_concat.py).  Other K: the scalar strided kernel loses to
        # "copy the slice, then the bulk-copy backward", so that is what runs.
        kq = K >> 2
        in_place = K % 4 == 0 and kq & (kq - 1) == 0 and kq <= 32 and all(
            c % 4 == 0 and o % 4 == 0 for c, o in zip(ctx.cstrides, ctx.offsets))
        if in_place:
            T = check(lib.gg_splat_bwd_tiles_concat(K, len(gb), sizes, int_array(hm), int_array([0] * len(gb)), strides,
                                                    offs), "gg_splat_bwd_tiles_concat")
            partials = torch.empty((B, T, K, GG_MOMENTS), dtype=torch.float32, device=dev)
            check(lib.gg_splat_bwd_concat(params.data_ptr(), B, K, len(gb), sizes, ptr_array([_ptr(g) for g in gb]),
                                          strides, offs, None, partials.data_ptr(), T, dev.index, _stream(dev)),
                  "gg_splat_bwd_concat")
        else:
            dense = [None if g is None else g[..., o:o + K].contiguous() for g, o in zip(gb, ctx.offsets)]
            partials, T = _splat_backward(params, B, K, spec, dense, [None] * len(dense))
        need = ctx.needs_input_grad
        dmus = torch.empty((B, K, 3), dtype=torch.float32, device=dev) if need[0] else None
        dsigma = torch.empty((B, K, 3, 3), dtype=torch.float64, device=dev) if need[1] else None
        drot = torch.empty((B, 3), dtype=torch.float32, device=dev) if any(need[2:5]) else None
        check(lib.gg_project_bwd(mus.data_ptr(), sigma.data_ptr(), int(sigma.dtype == torch.float64),
                                 rotx.data_ptr(), roty.data_ptr(), rotz.data_ptr(),
                                 spec.tx, spec.ty, spec.tz, spec.focal, B, K, partials.data_ptr(), T, None, None,
                                 _ptr(dmus), _ptr(dsigma), _ptr(drot), dev.index, _stream(dev)), "gg_project_bwd")
        if dsigma is not None and sigma.dtype != torch.float64:
            dsigma = dsigma.to(sigma.dtype)
        return (dmus, dsigma,
                drot[:, 0].contiguous() if need[2] else None,
                drot[:, 1].contiguous() if need[3] else None,
                drot[:, 2].contiguous() if need[4] else None,
                None, None) + tuple(gbufs)        # the other channels' gradients pass through untouched


# ----------------------------------------------------------------------------------- public API
def _prep_inputs(mus, sigma, rotx, roty, rotz, K):
    assert mus is not None                       # mask/mask_gen_dynamic.py:234
    assert sigma is not None                     # mask/mask_gen_dynamic.py:235
    dev = _require_cuda(mus, sigma, rotx, roty, rotz)
    if sigma.dim() != 4 or sigma.shape[-2:] != (3, 3):
        raise ValueError("sigma must be [B,K,3,3]")
    B, Kin = sigma.shape[0], sigma.shape[1]
    if K is not None and K != Kin:
        raise ValueError(f"K={K} does not match sigma.shape[1]={Kin}")
    if mus.numel() != B * Kin * 3:
        raise ValueError("mus must be [B,K,1,3] (reference) or [B,K,3]")
    mus3 = mus.reshape(B, Kin, 3)
    if mus3.dtype != torch.float32:
        mus3 = mus3.float()                      # reference: mus is float32, cast to f64 inside (:277)
    if sigma.dtype not in (torch.float32, torch.float64):
        sigma = sigma.double()
    rots = []
    for r in (rotx, roty, rotz):
        if r.numel() != B:
            raise ValueError("rotx/roty/rotz must be [B,1] (reference) or [B], in degrees")
        rots.append(r.reshape(B).float().contiguous())
    return dev, B, Kin, mus3.contiguous(), sigma.contiguous(), rots


def get_landmarks(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1., *, K: Optional[int] = None,
                  sizes: Sequence[int] = REFERENCE_SIZES, layout: str = "nhwc") -> List[torch.Tensor]:
    """Drop-in for the reference's ``get_landmarks`` (mask/mask_gen_dynamic.py:233-323).

    mus ``[B,K,1,3]`` f32, sigma ``[B,K,3,3]`` f64 (f32 accepted), rot* ``[B,1]`` f32 DEGREES,
    tx/ty/tz/focal Python floats.  Returns the list of ``[B,n,n,K]`` float32 maps for n in ``sizes``
    (reference: 32, 64, 128, 256).  ``K`` replaces the reference's module global ``nb_landmarks`` and is
    only checked against the tensors.  ``layout='nchw'`` returns ``[B,K,n,n]`` instead.
    """
    dev, B, Kk, mus3, sig, rots = _prep_inputs(mus, sigma, rotx, roty, rotz, K)
    sizes = tuple(int(s) for s in sizes)
    out: List[torch.Tensor] = []
    for i in range(0, len(sizes), _lib.GG_MAX_SCALES):       # the ABI takes <= 4 scales per launch
        chunk = sizes[i:i + _lib.GG_MAX_SCALES]
        spec = _Spec(float(tx), float(ty), float(tz), float(focal), chunk, (True,) * len(chunk),
                     (False,) * len(chunk), _layout_code(layout), False)
        out += list(_Render.apply(mus3, sig, rots[0], rots[1], rots[2], spec))
    return out


def get_landmarks_views(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1., *, K: Optional[int] = None,
                        sizes: Sequence[int] = REFERENCE_SIZES, layout: str = "nhwc",
                        with_mask: bool = False):
    """V cameras on the same Gaussians in ONE projection launch and ONE splat launch per direction.

    The reference calls ``get_landmarks`` three times per training step on the same ``(mus, sigmas)`` -- pose,
    pose + random yaw, canonical (mask/mask_gen_dynamic.py:707-709); at its batch size of 3 every call is
    launch-bound.  ``rotx, roty, rotz``: ``[V,B,1]`` (or ``[V,B]``) float32 degrees.  Returns a list over the V views of
    what ``get_landmarks`` returns for that view (a list of ``[B,n,n,K]`` maps, views of one ``[V*B,n,n,K]`` tensor per
    scale); with ``with_mask=True`` each entry is ``(maps_list, mask [B,n_last,n_last,1])`` as from
    ``get_landmarks_with_mask``.  Gradients of ``mus`` / ``sigma`` are accumulated over the views inside the projection
    adjoint; each view's angles get their own."""
    assert mus is not None and sigma is not None
    if sigma.dim() != 4 or sigma.shape[-2:] != (3, 3):
        raise ValueError("sigma must be [B,K,3,3]")
    B = sigma.shape[0]
    rx = torch.as_tensor(rotx)
    if rx.dim() < 2 or rx.numel() % B != 0:
        raise ValueError("rotx/roty/rotz must be [V,B,1] or [V,B]")
    V = rx.numel() // B
    flat = [torch.as_tensor(r).reshape(V * B) for r in (rotx, roty, rotz)]
    dev, _, Kk, mus3, sig, _ = _prep_inputs(mus, sigma, flat[0][:B], flat[1][:B], flat[2][:B], K)
    rots = [r.float().contiguous() for r in flat]
    _require_cuda(*rots)
    sizes = tuple(int(s) for s in sizes)
    if len(sizes) > _lib.GG_MAX_SCALES:
        raise ValueError("at most %d scales" % _lib.GG_MAX_SCALES)
    wk = tuple(with_mask and i == len(sizes) - 1 for i in range(len(sizes)))
    spec = _Spec(float(tx), float(ty), float(tz), float(focal), sizes, (True,) * len(sizes), wk, _layout_code(layout),
                 False, V)
    outs = _Render.apply(mus3, sig, rots[0], rots[1], rots[2], spec)
    maps = outs[:len(sizes)]
    res = []
    for v in range(V):
        mv = [m[v * B:(v + 1) * B] for m in maps]
        res.append((mv, outs[-1][v * B:(v + 1) * B].unsqueeze(-1)) if with_mask else mv)
    return res


def get_landmarks_into(bufs: Sequence[torch.Tensor], offsets: Sequence[int], mus, sigma, rotx, roty, rotz, tx, ty, tz,
                       focal=1., *, K: Optional[int] = None) -> List[torch.Tensor]:
    """``get_landmarks`` writing straight into the generator's concat buffers (concat-direct).

    The reference does ``x = tf.concat([x, pe[i]], -1)`` at every scale (mask/mask_gen_dynamic.py:523-529,
    rgb/rgb_gen_dynamic.py:604-610): a dense ``[B,n,n,K]`` tensor is written, read again and written a second time.
    Here ``bufs[i]`` is the pre-allocated NHWC concat buffer ``[B,n_i,n_i,C_i]`` (float32, contiguous; the caller has
    already put -- or will put -- ``x`` into its other channels) and the K maps of scale i land in channels
    ``[offsets[i], offsets[i] + K)`` in one pass.  The backward reads the same channel slice of the buffer's gradient
    in place.  Returns the buffers (autograd-connected: use the returned tensors downstream).  Scales are taken from
    the buffers' shapes; at most 4 per call."""
    dev, B, Kk, mus3, sig, rots = _prep_inputs(mus, sigma, rotx, roty, rotz, K)
    bufs = list(bufs)
    if not (1 <= len(bufs) <= _lib.GG_MAX_SCALES) or len(offsets) != len(bufs):
        raise ValueError("1..%d buffers with one channel offset each" % _lib.GG_MAX_SCALES)
    sizes = []
    for b, c0 in zip(bufs, offsets):
        _require_cuda(b)
        if b.dtype != torch.float32 or b.dim() != 4 or b.shape[0] != B or b.shape[1] != b.shape[2] or not b.is_contiguous():
            raise ValueError("each buffer must be a contiguous float32 NHWC tensor [B,n,n,C]")
        if c0 < 0 or c0 + Kk > b.shape[-1]:
            raise ValueError(f"channels [{c0}, {c0 + Kk}) do not fit a buffer with {b.shape[-1]} channels")
        sizes.append(int(b.shape[1]))
    spec = _Spec(float(tx), float(ty), float(tz), float(focal), tuple(sizes), (True,) * len(sizes),
                 (False,) * len(sizes), LAYOUT_NHWC, False)
    return list(_RenderInto.apply(mus3, sig, rots[0], rots[1], rots[2], spec, tuple(int(c) for c in offsets), *bufs))


def get_landmarks_with_mask(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1., *, K: Optional[int] = None,
                            sizes: Sequence[int] = REFERENCE_SIZES, layout: str = "nhwc"):
    """``get_landmarks`` plus the sum composite of the LAST scale (mask/mask_gen_dynamic.py:785-786),
    produced by the same launch.  Returns ``(maps_list, mask [B,n_last,n_last,1])``."""
    dev, B, Kk, mus3, sig, rots = _prep_inputs(mus, sigma, rotx, roty, rotz, K)
    sizes = tuple(int(s) for s in sizes)
    if len(sizes) > _lib.GG_MAX_SCALES:
        raise ValueError("at most %d scales" % _lib.GG_MAX_SCALES)
    wk = tuple(i == len(sizes) - 1 for i in range(len(sizes)))
    spec = _Spec(float(tx), float(ty), float(tz), float(focal), sizes, (True,) * len(sizes), wk,
                 _layout_code(layout), False)
    outs = _Render.apply(mus3, sig, rots[0], rots[1], rots[2], spec)
    return list(outs[:-1]), outs[-1].unsqueeze(-1)


def render_silhouette(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1., *, size: int = 256,
                      K: Optional[int] = None) -> torch.Tensor:
    """``reduce_sum(get_landmarks(...)[-1], axis=-1)`` (mask/mask_gen_dynamic.py:785-786) without ever
    writing the K-channel maps: returns the unclamped soft silhouette ``[B,size,size]`` float32."""
    dev, B, Kk, mus3, sig, rots = _prep_inputs(mus, sigma, rotx, roty, rotz, K)
    spec = _Spec(float(tx), float(ty), float(tz), float(focal), (int(size),), (False,), (True,), LAYOUT_NHWC, False)
    return _Render.apply(mus3, sig, rots[0], rots[1], rots[2], spec)[0]


def project(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1., *, K: Optional[int] = None):
    """The ``mu2d`` / ``sigma2d`` tensors of the reference graph (mask/mask_gen_dynamic.py:313-314), float64:
    ``([B,K,2], [B,K,2,2])``.  These are what the frozen decoder is fed (mask/infer_masks.py:394-395)."""
    dev, B, Kk, mus3, sig, rots = _prep_inputs(mus, sigma, rotx, roty, rotz, K)
    spec = _Spec(float(tx), float(ty), float(tz), float(focal), (), (), (), LAYOUT_NHWC, True)
    mu2d, sigma2d = _Render.apply(mus3, sig, rots[0], rots[1], rots[2], spec)
    return mu2d, sigma2d


def get_gaussian_maps_2d(mu, sigma, shape_hw, mode: str = "rot", *, layout: str = "nhwc") -> torch.Tensor:
    """Drop-in for the reference's ``get_gaussian_maps_2d`` (mask/mask_gen_dynamic.py:112-143).

    mu ``[B,K,2]`` (x = column, y = row), sigma ``[B,K,2,2]`` -> ``[B,H,W,K]``.  Like the reference
    (whose tile/BatchMatMul construction only type-checks for square maps, :133-136) H must equal W.
    The result is float32 (the reference casts right after the call, :318).
    """
    if mode != "rot":
        raise ValueError("Unknown mode: " + str(mode))        # mirrors :189-190 of the dead sibling
    _require_cuda(mu, sigma)
    H, W = int(shape_hw[0]), int(shape_hw[1])
    if H != W:
        raise ValueError("get_gaussian_maps_2d requires H == W (reference tiles invsigma over W, :134)")
    if sigma.dim() != 4 or sigma.shape[-2:] != (2, 2):
        raise ValueError("sigma must be [B,K,2,2]")
    B, K = sigma.shape[0], sigma.shape[1]
    mu = mu.reshape(B, K, 2)
    dt = torch.float64 if (mu.dtype == torch.float64 or sigma.dtype == torch.float64) else torch.float32
    spec = _Spec(0., 0., 0., 1., (H,), (True,), (False,), _layout_code(layout), False)
    return _Raster.apply(mu.to(dt).contiguous(), sigma.to(dt).contiguous(), spec)[0]


def get_silhouette_2d(mu, sigma, shape_hw) -> torch.Tensor:
    """``reduce_sum(get_gaussian_maps_2d(mu, sigma, shape_hw), -1)`` (mask/mask_gen_dynamic.py:112-143 + :785-786) from
    the frozen-graph boundary tensors ``mu2d [B,K,2]`` / ``sigma2d [B,K,2,2]`` (mask/infer_masks.py:394-395), without the
    K-channel maps: the mask-only kernels (forward-differenced by default), differentiable w.r.t. both inputs."""
    _require_cuda(mu, sigma)
    H, W = int(shape_hw[0]), int(shape_hw[1])
    if H != W:
        raise ValueError("get_silhouette_2d requires H == W (reference tiles invsigma over W, :134)")
    if sigma.dim() != 4 or sigma.shape[-2:] != (2, 2):
        raise ValueError("sigma must be [B,K,2,2]")
    B, K = sigma.shape[0], sigma.shape[1]
    mu = mu.reshape(B, K, 2)
    dt = torch.float64 if (mu.dtype == torch.float64 or sigma.dtype == torch.float64) else torch.float32
    spec = _Spec(0., 0., 0., 1., (H,), (False,), (True,), LAYOUT_NHWC, False)
    return _Raster.apply(mu.to(dt).contiguous(), sigma.to(dt).contiguous(), spec)[0]


def get_landmarks_infer(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal=1., *, size: int = 256,
                        K: Optional[int] = None):
    """Inference flavour (mask/infer_masks.py:186-267): float64 projection, ``mu2d``/``sigma2d`` cast to
    float32, then a float32 raster at 256 only.  Returns ``(mu2d, sigma2d, maps)``."""
    mu2d, sigma2d = project(mus, sigma, rotx, roty, rotz, tx, ty, tz, focal, K=K)
    mu2d = mu2d.float()
    sigma2d = sigma2d.float()
    return mu2d, sigma2d, get_gaussian_maps_2d(mu2d, sigma2d, [size, size])
