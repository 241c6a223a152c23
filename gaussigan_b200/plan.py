"""Pre-planned silhouette step: the renderer's forward + backward with every buffer allocated once.

``torch.autograd`` costs tens of microseconds of host time per call -- comparable with the ~50 us the
whole B=64 step takes on a B200 -- so callers that run the path in a loop (training, the benchmark)
can plan it once:

    step = SilhouetteStep(B, K, size, device)
    mask = step.forward(mus, sigma, rotx, roty, rotz)          # [B,size,size] f32 (sum composite)
    dmus, dsigma, drot = step.backward(gmask)                  # grads into Gaussians and camera

Both methods are plain sequences of C-ABI launches on the current stream (capturable in a CUDA
graph: no allocation, no host synchronisation).  ``capture()`` records forward+backward into one
graph.  Semantics are identical to ``render_silhouette(...)`` + ``.backward(gmask)``
(mask/mask_gen_dynamic.py:233-323, 785-786 and their tf.gradients).
"""
from __future__ import annotations

from typing import Optional

import torch

from ._lib import GG_MOMENTS, GG_PARAM_STRIDE, LAYOUT_NHWC, check, int_array, lib, ptr_array


class SilhouetteStep:
    def __init__(self, B: int, K: int, size: int = 256, device: torch.device | str = "cuda",
                 tx: float = 0.0, ty: float = 0.0, tz: float = -2.0, focal: float = 1.0,
                 sigma_dtype: torch.dtype = torch.float64):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("gaussigan_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        self.B, self.K, self.n, self.dev = int(B), int(K), int(size), dev
        self.t = (float(tx), float(ty), float(tz))
        self.focal = float(focal)
        self.sigma_f64 = int(sigma_dtype == torch.float64)
        f32, f64 = torch.float32, torch.float64
        self.params = torch.empty((B, K, GG_PARAM_STRIDE), dtype=f32, device=dev)
        self.mask = torch.empty((B, size, size), dtype=f32, device=dev)
        self._sizes = int_array([size])
        self._one = int_array([1])
        self._zero = int_array([0])
        self.T = check(lib.gg_splat_bwd_tiles(K, 1, self._sizes, self._zero, self._one, LAYOUT_NHWC),
                       "gg_splat_bwd_tiles")
        self.partials = torch.empty((B, self.T, K, GG_MOMENTS), dtype=f32, device=dev)
        self.dmus = torch.empty((B, K, 3), dtype=f32, device=dev)
        self.dsigma = torch.empty((B, K, 3, 3), dtype=f64, device=dev)
        self.drot = torch.empty((B, 3), dtype=f32, device=dev)
        self._mask_ptrs = ptr_array([self.mask.data_ptr()])
        self._null_ptrs = ptr_array([None])
        self._inputs = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    # ------------------------------------------------------------------ plain launches
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.dev).cuda_stream

    def bind(self, mus, sigma, rotx, roty, rotz):
        """Remember the (static) input tensors; mus [B,K,3] f32, sigma [B,K,3,3], rot* [B] f32 deg."""
        B, K = self.B, self.K
        assert mus.is_cuda and mus.dtype == torch.float32 and mus.numel() == B * K * 3 and mus.is_contiguous()
        assert sigma.is_contiguous() and sigma.numel() == B * K * 9
        assert int(sigma.dtype == torch.float64) == self.sigma_f64
        for r in (rotx, roty, rotz):
            assert r.dtype == torch.float32 and r.numel() == B and r.is_contiguous()
        self._inputs = (mus, sigma, rotx, roty, rotz)

    def forward(self, mus=None, sigma=None, rotx=None, roty=None, rotz=None) -> torch.Tensor:
        if mus is not None:
            self.bind(mus, sigma, rotx, roty, rotz)
        mus, sigma, rotx, roty, rotz = self._inputs
        st = self._stream()
        check(lib.gg_project_fwd(mus.data_ptr(), sigma.data_ptr(), self.sigma_f64, rotx.data_ptr(),
                                 roty.data_ptr(), rotz.data_ptr(), *self.t, self.focal, self.B, self.K,
                                 None, None, self.params.data_ptr(), self.dev.index, st), "gg_project_fwd")
        check(lib.gg_splat_fwd(self.params.data_ptr(), self.B, self.K, 1, self._sizes, self._null_ptrs,
                               self._mask_ptrs, LAYOUT_NHWC, self.dev.index, st), "gg_splat_fwd")
        return self.mask

    def backward(self, gmask: torch.Tensor):
        """gmask [B,n,n] f32 contiguous = dL/dmask -> (dmus [B,K,3] f32, dsigma [B,K,3,3] f64, drot [B,3] f32)."""
        assert gmask.is_cuda and gmask.dtype == torch.float32 and gmask.is_contiguous()
        assert gmask.numel() == self.B * self.n * self.n
        mus, sigma, rotx, roty, rotz = self._inputs
        st = self._stream()
        check(lib.gg_splat_bwd(self.params.data_ptr(), self.B, self.K, 1, self._sizes, self._null_ptrs,
                               ptr_array([gmask.data_ptr()]), LAYOUT_NHWC, self.partials.data_ptr(), self.T,
                               self.dev.index, st), "gg_splat_bwd")
        check(lib.gg_project_bwd(mus.data_ptr(), sigma.data_ptr(), self.sigma_f64, rotx.data_ptr(),
                                 roty.data_ptr(), rotz.data_ptr(), *self.t, self.focal, self.B, self.K,
                                 self.partials.data_ptr(), self.T, None, None, self.dmus.data_ptr(),
                                 self.dsigma.data_ptr(), self.drot.data_ptr(), self.dev.index, st),
              "gg_project_bwd")
        return self.dmus, self.dsigma, self.drot

    #: kernels launched by one forward() + backward()
    LAUNCHES_PER_STEP = 4

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, gmask: torch.Tensor) -> torch.cuda.CUDAGraph:
        """Record forward() + backward(gmask) on the bound inputs into one CUDA graph."""
        assert self._inputs is not None, "bind() inputs first"
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.forward()
            self.backward(gmask)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.forward()
            self.backward(gmask)
        self.graph = g
        return g


class SilhouetteLossStep(SilhouetteStep):
    """``SilhouetteStep`` with the consumer fused in: projection -> ONE kernel that renders the silhouette,
    evaluates ``mean|0.5*(target+1) - silhouette|`` (mask/mask_gen_dynamic.py:785-787), forms its gradient and
    runs the splat backward -> projection adjoint.  The silhouette and dL/dmask never reach HBM.  K <= 64.

        step = SilhouetteLossStep(B, K, size, device)
        step.bind(mus, sigma, rotx, roty, rotz)
        dmus, dsigma, drot = step.loss_backward(target)     # target [B,n,n] uint8 (raw 0..255) or f32 in [-1,1]
        step.loss()                                          # float (synchronises)
    """

    LAUNCHES_PER_STEP = 3

    def __init__(self, *a, want_mask: bool = False, **kw):
        super().__init__(*a, **kw)
        if self.K > 64:
            raise ValueError("the fused loss kernel handles K <= 64")
        self.loss_partials = torch.empty(self.B * self.T, dtype=torch.float32, device=self.dev)
        self.want_mask = want_mask

    def loss_backward(self, target: torch.Tensor, bits: bool = False, strips: bool = False):
        """``target`` [B,n,n] uint8 (raw 0..255) or float32 in [-1,1]; with ``bits=True`` a two-level mask packed one
        bit per pixel, uint8 [B,n,ceil(n/8)] (``numpy.packbits(mask > 0, axis=-1, bitorder="little")``); with
        ``strips=True`` the device copy of a ``gaussigan_b200.host.encode_target_strips`` buffer (any 8-bit mask, lossless)."""
        assert target.is_cuda and target.is_contiguous()
        if strips:
            assert target.dtype == torch.uint8 and target.data_ptr() % 8 == 0
            kind = 3
        elif bits:
            assert target.dtype == torch.uint8 and target.numel() == self.B * self.n * ((self.n + 7) // 8)
            kind = 2
        else:
            assert target.numel() == self.B * self.n * self.n and target.dtype in (torch.uint8, torch.float32)
            kind = 0 if target.dtype == torch.uint8 else 1
        mus, sigma, rotx, roty, rotz = self._inputs
        st = self._stream()
        check(lib.gg_project_fwd(mus.data_ptr(), sigma.data_ptr(), self.sigma_f64, rotx.data_ptr(),
                                 roty.data_ptr(), rotz.data_ptr(), *self.t, self.focal, self.B, self.K,
                                 None, None, self.params.data_ptr(), self.dev.index, st), "gg_project_fwd")
        check(lib.gg_silhouette_loss_fused(self.params.data_ptr(), target.data_ptr(), kind, self.B, self.K, self.n,
                                           self.mask.data_ptr() if self.want_mask else None,
                                           self.partials.data_ptr(), self.T, self.loss_partials.data_ptr(),
                                           self.dev.index, st), "gg_silhouette_loss_fused")
        check(lib.gg_project_bwd(mus.data_ptr(), sigma.data_ptr(), self.sigma_f64, rotx.data_ptr(),
                                 roty.data_ptr(), rotz.data_ptr(), *self.t, self.focal, self.B, self.K,
                                 self.partials.data_ptr(), self.T, None, None, self.dmus.data_ptr(),
                                 self.dsigma.data_ptr(), self.drot.data_ptr(), self.dev.index, st),
              "gg_project_bwd")
        return self.dmus, self.dsigma, self.drot

    def loss(self) -> float:
        return float(self.loss_partials.double().sum().item() / (self.B * self.n * self.n))

    def capture_loss(self, target: torch.Tensor) -> torch.cuda.CUDAGraph:
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.loss_backward(target)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.loss_backward(target)
        self.graph = g
        return g


class DynamicObjectStep:
    """Config 3 of BASELINE.json planned once: the dynamic-object mask step of the reference
    (mask/mask_gen_dynamic.py:383-401 frame -> canonical covariances, :430 SVD, :426-458 per-frame deformation,
    :707 render, :785-787 ``m_recon_loss_bis``) and its whole backward, as a fixed sequence of C-ABI launches on
    pre-allocated buffers -- no autograd bookkeeping, no host synchronisation, capturable in one CUDA graph.

        step = DynamicObjectStep(B, K, size, device)
        step.bind(v0, v1, u_raw, cannmu, mu_t, scale, ang, theta_raw)    # static input tensors (head OUTPUTS, after tanh)
        step.loss_backward(target)                                       # target [B,n,n] uint8 | float32 in [-1,1]
        step.loss(); step.grads                                          # dict of gradients w.r.t. every bound input

    Shapes: ``v0, v1, u_raw [K,3]`` f32 (canonical frame heads), ``cannmu [K,3]`` f32, ``mu_t, scale, ang [B,K,3]``
    f32 (``ang`` = per-Gaussian Euler x,y,z), ``theta_raw [B]`` f32.  Gradients of the batch-shared canonical
    parameters are the sums over the batch (the broadcast of :428, :455).
    """

    def __init__(self, B: int, K: int, size: int = 256, device="cuda", tx=0.0, ty=0.0, tz=-2.0, focal=1.0):
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("gaussigan_b200 runs on CUDA (sm_100a) only; there is no CPU path")
        if dev.index is None:
            dev = torch.device("cuda", torch.cuda.current_device())
        if K > 64:
            raise ValueError("the fused loss kernel handles K <= 64")
        self.B, self.K, self.n, self.dev = int(B), int(K), int(size), dev
        self.t, self.focal = (float(tx), float(ty), float(tz)), float(focal)
        f32 = dict(dtype=torch.float32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        e = torch.empty
        # forward intermediates
        self.csig, self.U, self.V, self.S = e((K, 3, 3), **f64), e((K, 3, 3), **f64), e((K, 3, 3), **f64), e((K, 3), **f64)
        self.mus, self.sig, self.theta = e((B, K, 3), **f32), e((B, K, 3, 3), **f64), e((B,), **f32)
        self.zeros = torch.zeros((B,), **f32)                      # rotx = rotz = 0 (:706-709)
        self.params = e((B, K, GG_PARAM_STRIDE), **f32)
        self._sizes, self._one, self._zero = int_array([size]), int_array([1]), int_array([0])
        self.T = check(lib.gg_splat_bwd_tiles(K, 1, self._sizes, self._zero, self._one, LAYOUT_NHWC), "gg_splat_bwd_tiles")
        self.partials = e((B, self.T, K, GG_MOMENTS), **f32)
        self.loss_partials = e((B * self.T,), **f32)
        # backward intermediates
        self.dmus, self.dsig, self.drot = e((B, K, 3), **f32), e((B, K, 3, 3), **f64), e((B, 3), **f32)
        self.gtheta = e((B,), **f32)
        self.dcmu_b, self.dU_b, self.dS_b = e((B, K, 3), **f32), e((B, K, 3, 3), **f64), e((B, K, 3), **f64)
        self.dU, self.dS, self.gcsig = e((K, 3, 3), **f64), e((K, 3), **f64), e((K, 3, 3), **f64)
        self.grads = {"v0": e((K, 3), **f32), "v1": e((K, 3), **f32), "u_raw": e((K, 3), **f32), "cannmu": e((K, 3), **f32),
                      "mu_t": e((B, K, 3), **f32), "scale": e((B, K, 3), **f32), "ang": e((B, K, 3), **f32),
                      "theta_raw": e((B,), **f32)}
        self._in = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    #: C-ABI launches per step (plus four tiny torch reductions / copies)
    LAUNCHES_PER_STEP = 10

    def bind(self, v0, v1, u_raw, cannmu, mu_t, scale, ang, theta_raw):
        B, K = self.B, self.K
        for t, shape in ((v0, (K, 3)), (v1, (K, 3)), (u_raw, (K, 3)), (cannmu, (K, 3)), (mu_t, (B, K, 3)),
                         (scale, (B, K, 3)), (ang, (B, K, 3)), (theta_raw, (B,))):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape, (t.shape, shape)
        self._in = (v0, v1, u_raw, cannmu, mu_t, scale, ang, theta_raw)

    def loss_backward(self, target: torch.Tensor):
        assert self._in is not None, "bind() inputs first"
        assert target.is_cuda and target.is_contiguous() and target.numel() == self.B * self.n * self.n
        assert target.dtype in (torch.uint8, torch.float32)
        v0, v1, u_raw, cannmu, mu_t, scale, ang, theta_raw = self._in
        B, K, n, d = self.B, self.K, self.n, self.dev.index
        st = torch.cuda.current_stream(self.dev).cuda_stream
        p = lambda t: t.data_ptr()
        g = self.grads
        # ---- forward
        check(lib.gg_sigma_from_frame_fwd(p(v0), p(v1), p(u_raw), K, p(self.csig), d, st), "gg_sigma_from_frame_fwd")
        check(lib.gg_svd3_fwd(p(self.csig), K, p(self.U), p(self.S), p(self.V), d, st), "gg_svd3_fwd")
        check(lib.gg_deform_fwd(p(mu_t), p(scale), p(ang), p(cannmu), p(self.U), p(self.S), p(theta_raw), B, K,
                                p(self.mus), p(self.sig), p(self.theta), d, st), "gg_deform_fwd")
        check(lib.gg_project_fwd(p(self.mus), p(self.sig), 1, p(self.zeros), p(self.theta), p(self.zeros), *self.t,
                                 self.focal, B, K, None, None, p(self.params), d, st), "gg_project_fwd")
        check(lib.gg_silhouette_loss_fused(p(self.params), p(target), 0 if target.dtype == torch.uint8 else 1, B, K, n,
                                           None, p(self.partials), self.T, p(self.loss_partials), d, st),
              "gg_silhouette_loss_fused")
        # ---- backward
        check(lib.gg_project_bwd(p(self.mus), p(self.sig), 1, p(self.zeros), p(self.theta), p(self.zeros), *self.t,
                                 self.focal, B, K, p(self.partials), self.T, None, None, p(self.dmus), p(self.dsig),
                                 p(self.drot), d, st), "gg_project_bwd")
        self.gtheta.copy_(self.drot[:, 1])                          # d loss / d roty
        check(lib.gg_deform_bwd(p(scale), p(ang), p(self.U), p(self.S), p(self.dmus), p(self.dsig), p(self.gtheta), B, K,
                                p(g["theta_raw"]), p(g["mu_t"]), p(g["scale"]), p(g["ang"]), p(self.dcmu_b),
                                p(self.dU_b), p(self.dS_b), d, st), "gg_deform_bwd")
        torch.sum(self.dcmu_b, dim=0, out=g["cannmu"])              # canonical parameters are broadcast over the batch
        torch.sum(self.dU_b, dim=0, out=self.dU)
        torch.sum(self.dS_b, dim=0, out=self.dS)
        check(lib.gg_svd3_bwd(p(self.U), p(self.S), p(self.V), p(self.dU), p(self.dS), K, p(self.gcsig), d, st),
              "gg_svd3_bwd")
        check(lib.gg_sigma_from_frame_bwd(p(v0), p(v1), p(u_raw), p(self.gcsig), K, p(g["v0"]), p(g["v1"]),
                                          p(g["u_raw"]), d, st), "gg_sigma_from_frame_bwd")
        return g

    def loss(self) -> float:
        return float(self.loss_partials.double().sum().item() / (self.B * self.n * self.n))

    def capture(self, target: torch.Tensor) -> torch.cuda.CUDAGraph:
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            self.loss_backward(target)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            self.loss_backward(target)
        self.graph = gr
        return gr
