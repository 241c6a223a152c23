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

    def loss_backward(self, target: torch.Tensor):
        assert target.is_cuda and target.is_contiguous() and target.numel() == self.B * self.n * self.n
        assert target.dtype in (torch.uint8, torch.float32)
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
