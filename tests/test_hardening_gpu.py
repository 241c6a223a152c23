"""Round-2 parity hardening (VERDICT r1 "harden parity"): pointwise statistics of both arithmetic modes, deterministic
cases on the guard's three boundaries, the K range the ABI advertises, and the reference-side stub -- all through the
C ABI on the GPU, against the float64 oracle."""
import importlib.util
import math
import os

import numpy as np
import pytest
import torch

from oracle import reference_port as rp
from util import FWD_RTOL, GRAD_RTOL, assert_close_norm, max_rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG2E = 1.4426950408889634


@pytest.fixture(scope="module")
def gg(cuda_device):
    import gaussigan_b200
    return gaussigan_b200


@pytest.fixture
def restore_modes(gg):
    yield
    gg.set_fast_mask(True)
    gg.set_pdl(True)


# ------------------------------------------------------------------------------------ pointwise statistics at C2
def test_pointwise_allclose_statistics_at_config2(gg, cuda_device, restore_modes):
    """SURVEY.md 7.2 offers two formulations of the 1e-5 forward tolerance: norm-relative |d| <= 1e-5 max|ref| (the one
    tests/util.py uses) and pointwise allclose(rtol=1e-5, atol=1e-6).  This test measures the second one at the full C2
    shape (batch 64, K=16, 256x256) for both arithmetic modes of the mask-only kernels and bounds it:
      direct form (one exponential per pixel-Gaussian): no pixel violates it;
      forward-differenced form (default): absolute error <= 1e-5 everywhere and <= 0.5 % of the pixels outside the
      pointwise band -- it meets the norm-relative formulation, not the pointwise one, on a silhouette
      whose values reach ~6."""
    B, K, n = 64, 16, 256
    inp = rp.synth_inputs(B, K, seed=56)
    ref = torch.cat([rp.get_landmarks(inp["mus"][i:i + 16], inp["sigma"][i:i + 16], inp["rotx"][i:i + 16],
                                      inp["roty"][i:i + 16], inp["rotz"][i:i + 16], 0, 0, -2., sizes=(n,))[0].sum(-1)
                     for i in range(0, B, 16)]).double()
    d = {k: v.to(cuda_device) for k, v in inp.items()}
    stats = {}
    for fast in (True, False):
        gg.set_fast_mask(fast)
        m = gg.render_silhouette(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., size=n).double().cpu()
        err = (m - ref).abs()
        viol = err > (1e-6 + 1e-5 * ref.abs())
        stats[fast] = dict(max_abs=float(err.max()), norm_rel=float(err.max() / ref.abs().max()),
                           viol_frac=float(viol.double().mean()),
                           worst_pointwise=float((err / (1e-6 + 1e-5 * ref.abs())).max()))
    print("pointwise statistics (fast, direct):", stats)
    assert stats[False]["viol_frac"] == 0.0 and stats[False]["norm_rel"] <= 1e-6
    assert stats[True]["norm_rel"] <= FWD_RTOL / 2 and stats[True]["max_abs"] <= 1e-5
    assert stats[True]["viol_frac"] <= 5e-3 and stats[True]["worst_pointwise"] <= 5.0


# ------------------------------------------------------------------------------------------ the guard's boundaries
def _guard_case(n, nA_abs, tmax_pixels, shear, B=1):
    """2-D Gaussians (mu2d, sigma2d) with a prescribed sharpness |nA| = a log2(e) and centre offset, one per batch
    element, placed so that the strip recurrence's three criteria (gg_fast.cuh: |nA| <= 512, visible slope
    2 h sqrt(14 |nA|) <= 3, range slope |2 nA h| (tmax + 8 h) <= 10) sit just inside / just outside."""
    h = 2.0 / (n - 1)
    a = nA_abs / LOG2E                       # conic entry P00
    c = a * 0.05 + 0.3                       # a broad axis along y
    b = shear * math.sqrt(a * c)
    P = torch.tensor([[a, b], [b, c]], dtype=torch.float64)
    sigma = torch.linalg.inv(P).reshape(1, 1, 2, 2).repeat(B, 1, 1, 1)
    mu = torch.tensor([[[-1.0 + tmax_pixels * h * (0.5 if tmax_pixels < n else 1.0), 0.07]]], dtype=torch.float64).repeat(B, 1, 1)
    return mu, sigma


def _check_2d(gg, dev, mu, sigma, n, what):
    mu_r, sg_r = mu.clone().requires_grad_(True), sigma.clone().requires_grad_(True)
    ref = rp.get_gaussian_maps_2d(mu_r, sg_r, [n, n]).sum(-1)
    gm = torch.randn(ref.shape, generator=torch.Generator().manual_seed(n), dtype=torch.float64)
    (ref * gm).sum().backward()
    mu_d, sg_d = mu.to(dev).requires_grad_(True), sigma.to(dev).requires_grad_(True)
    m = gg.get_silhouette_2d(mu_d, sg_d, [n, n])
    m.backward(gm.float().to(dev))
    assert torch.isfinite(m).all() and torch.isfinite(mu_d.grad).all() and torch.isfinite(sg_d.grad).all(), what
    if float(ref.detach().abs().max()) < 1e-30:       # so far off the frame that float32 holds nothing: exact zeros
        assert float(m.abs().max()) <= 1e-30 and float(mu_d.grad.abs().max()) <= 1e-25, what
        return 0.0, 0.0
    assert_close_norm(m, ref, FWD_RTOL, what + " forward")
    assert_close_norm(mu_d.grad, mu_r.grad, GRAD_RTOL, what + " d mu2d")
    assert_close_norm(sg_d.grad, sg_r.grad, GRAD_RTOL, what + " d sigma2d")
    return max_rel(m, ref), max(max_rel(mu_d.grad, mu_r.grad), max_rel(sg_d.grad, sg_r.grad))


@pytest.mark.parametrize("n", [256, 128, 64, 24, 20, 16, 13])
@pytest.mark.parametrize("eps", [-2e-3, 2e-3])
def test_guard_boundary_sharpness(gg, cuda_device, n, eps):
    """|nA| = 512 (1 +- eps): the recurrence / direct switch of criterion (iii)."""
    mu, sg = _guard_case(n, 512.0 * (1 + eps), tmax_pixels=n // 3, shear=0.3)
    _check_2d(gg, cuda_device, mu, sg, n, f"|nA|=512(1{eps:+g}) n={n}")


@pytest.mark.parametrize("n", [256, 128, 64, 24, 20, 16, 13])
@pytest.mark.parametrize("eps", [-1e-2, 1e-2])
def test_guard_boundary_visible_slope(gg, cuda_device, n, eps):
    """2 h sqrt(14 |nA|) = 3 (1 +- eps): criterion (ii), the per-pixel exponent change where the Gaussian is visible."""
    h = 2.0 / (n - 1)
    nA = min(((3.0 * (1 + eps)) / (2 * h)) ** 2 / 14.0, 511.0)       # stay below the sharpness limit so (ii) decides
    mu, sg = _guard_case(n, nA, tmax_pixels=n // 2, shear=-0.4)
    _check_2d(gg, cuda_device, mu, sg, n, f"visible slope 3(1{eps:+g}) n={n}")


@pytest.mark.parametrize("n", [256, 128, 64, 24, 16, 13])
@pytest.mark.parametrize("eps", [-1e-2, 1e-2])
def test_guard_boundary_range_slope(gg, cuda_device, n, eps):
    """|2 nA h| (tmax + 8 h) = 10 (1 +- eps): criterion (i), range safety, with the centre far off the tile."""
    h = 2.0 / (n - 1)
    nA = min(0.5 * ((3.0 / (2 * h)) ** 2 / 14.0), 400.0)              # comfortably inside (ii) and (iii)
    tmax = 10.0 * (1 + eps) / abs(2 * nA * h) - 8 * h                  # distance of the far tile corner from the centre
    # centre to the LEFT of the frame so that the far (right) edge is tmax away: mu_x = 1 - tmax (may be off-screen)
    a = nA / LOG2E
    P = torch.tensor([[a, 0.0], [0.0, 0.4]], dtype=torch.float64)
    sg = torch.linalg.inv(P).reshape(1, 1, 2, 2)
    mu = torch.tensor([[[1.0 - tmax, -0.2]]], dtype=torch.float64)
    _check_2d(gg, cuda_device, mu, sg, n, f"range slope 10(1{eps:+g}) n={n}")


@pytest.mark.parametrize("n", list(range(13, 25)))
def test_needles_on_coarse_grids(gg, cuda_device, n):
    """The cases the round-1 soaks found (n = 13 .. 24, a needle far narrower than a pixel next to ordinary blobs): the
    CENTRE variant of the backward sums guarded Gaussians about their own centre line."""
    B, K = 2, 5
    g = torch.Generator().manual_seed(1000 + n)
    inp = rp.synth_inputs(B, K, seed=300 + n, angle=40.0)
    inp["sigma"][0, 0] = torch.diag(torch.tensor([2e-4, 0.3, 0.04], dtype=torch.float64))
    inp["sigma"][1, 2] = torch.diag(torch.tensor([0.25, 3e-4, 0.05], dtype=torch.float64))
    lv = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    ref = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0, 0, -2., sizes=(n,))[0].sum(-1)
    gm = torch.randn(B, n, n, generator=g)
    (ref * gm).sum().backward()
    dv = {k: v.clone().to(cuda_device).requires_grad_(True) for k, v in inp.items()}
    m = gg.render_silhouette(dv["mus"], dv["sigma"], dv["rotx"], dv["roty"], dv["rotz"], 0, 0, -2., size=n)
    m.backward(gm.to(cuda_device))
    assert_close_norm(m, ref, FWD_RTOL, f"needles n={n} mask")
    for k in ("mus", "sigma", "roty"):
        assert_close_norm(dv[k].grad, lv[k].grad, GRAD_RTOL, f"needles n={n} grad {k}")


# ------------------------------------------------------------------------------------------- advertised K range
@pytest.mark.parametrize("K", [1025, 1536])
def test_mask_forward_accepts_the_whole_advertised_k_range(gg, cuda_device, restore_modes, K):
    """K in 1025 .. 1536 needs more than the default 48 KB of dynamic shared memory in the forward-differenced
    silhouette kernel: the launcher opts in; results equal the direct form and the oracle."""
    B, n = 1, 24
    inp = rp.synth_inputs(B, K, seed=K)
    ref = rp.get_landmarks(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0, 0, -2., sizes=(n,))[0].sum(-1)
    d = {k: v.to(cuda_device) for k, v in inp.items()}
    out = {}
    for fast in (True, False):
        gg.set_fast_mask(fast)
        out[fast] = gg.render_silhouette(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., size=n)
        assert_close_norm(out[fast], ref, FWD_RTOL, f"K={K} fast={fast}")


# ------------------------------------------------------------------------------------------ reference-side stub
def test_reference_side_stub_vs_oracle(cuda_device):
    """examples/gaussigan_binding.py (numpy + ctypes, no torch: what INTEGRATION.md tells a maintainer of the reference to
    add) against the oracle's loss and gradients."""
    spec = importlib.util.spec_from_file_location("gaussigan_binding", os.path.join(ROOT, "examples", "gaussigan_binding.py"))
    stub = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(stub)
    B, K, n = 5, 16, 64
    inp = rp.synth_inputs(B, K, seed=77)
    tgt = rp.synth_target_masks(B, n, seed=78)
    _, loss_ref, dmu_ref, dsig_ref, drot_ref = rp.render_loss_and_grads(inp, tgt, n)
    raw = ((tgt.reshape(B, n, n) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).numpy()
    loss, dmus, dsig, drot = stub.silhouette_loss_and_grads(inp["mus"].reshape(B, K, 3).numpy(), inp["sigma"].numpy(),
                                                            inp["rotx"].numpy(), inp["roty"].numpy(), inp["rotz"].numpy(), raw)
    assert abs(loss - float(loss_ref)) <= 1e-5 * abs(float(loss_ref))
    assert_close_norm(torch.from_numpy(dmus), dmu_ref.reshape(B, K, 3), GRAD_RTOL, "stub dmus")
    assert_close_norm(torch.from_numpy(dsig), dsig_ref, GRAD_RTOL, "stub dsigma")
    assert_close_norm(torch.from_numpy(drot[:, 1]), drot_ref.reshape(B), GRAD_RTOL, "stub droty")


def test_device_guard_restores_the_current_device(gg, cuda_device):
    """Every C-ABI entry makes its `device` argument current only for the duration of the call."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    inp = rp.synth_inputs(2, 4, seed=1)
    d1 = torch.device("cuda", 1)
    torch.cuda.set_device(0)
    dv = {k: v.to(d1) for k, v in inp.items()}
    m = gg.render_silhouette(dv["mus"], dv["sigma"], dv["rotx"], dv["roty"], dv["rotz"], 0, 0, -2., size=32)
    assert m.device == d1 and torch.cuda.current_device() == 0
    ref = rp.get_landmarks(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0, 0, -2., sizes=(32,))[0].sum(-1)
    assert_close_norm(m, ref, FWD_RTOL, "render on cuda:1 while cuda:0 is current")


# ------------------------------------------------------------------------------- concat-direct
@pytest.mark.parametrize("K,cx,sizes", [
    (16, 64, (32, 64, 128, 256)),      # quads, 128-bit accesses, the reference pyramid
    (16, 8, (64,)), (4, 4, (40,)),     # quads, small buffers
    (12, 32, (32, 64)),                # K % 4 == 0 but not 4 * 2^m: quadg forward, strip backward
    (6, 64, (64, 128)),                # the reference's K = 6: pair stores (64-bit), strip backward
    (6, 3, (48,)),                     # odd channel offset: scalar accesses
    (5, 7, (24,)),                     # odd everything
    (16, 6, (32,)),                    # quads of Gaussians but a channel offset that breaks 16-byte alignment
])
def test_concat_direct_forward_and_backward_vs_oracle(gg, cuda_device, K, cx, sizes):
    """get_landmarks_into: the K maps land in channels [cx, cx + K) of a pre-allocated NHWC concat buffer
    (mask/mask_gen_dynamic.py:523-529 `tf.concat([x, pe[i]], -1)`), and the backward reads that channel slice of the
    buffer's gradient in place.  Against torch.cat([x, oracle_maps], -1) and the oracle's autograd; the channels that
    belong to x are never touched; gradients reach x, mus, sigma and roty."""
    B = 2
    inp = rp.synth_inputs(B, K, seed=500 + K + cx, angle=30.0)
    g = torch.Generator().manual_seed(K * 100 + cx)
    xs = [torch.randn(B, n, n, cx, generator=g) for n in sizes]
    ws = [torch.randn(B, n, n, cx + K + 3, generator=g) for n in sizes]          # + 3 trailing channels nobody writes
    # ---- oracle
    lv = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    xr = [x.clone().requires_grad_(True) for x in xs]
    maps = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0, 0, -2., sizes=sizes)
    tail = [torch.full((B, n, n, 3), 7.0) for n in sizes]
    cat_ref = [torch.cat([x, m, t], -1) for x, m, t in zip(xr, maps, tail)]
    sum((c * w).sum() for c, w in zip(cat_ref, ws)).backward()
    # ---- concat-direct
    dv = {k: v.clone().to(cuda_device).requires_grad_(True) for k, v in inp.items()}
    xd = [x.clone().to(cuda_device).requires_grad_(True) for x in xs]
    bufs = []
    for x, n in zip(xd, sizes):
        b = torch.full((B, n, n, cx + K + 3), 7.0, device=cuda_device)
        b[..., :cx] = x
        bufs.append(b)
    outs = gg.get_landmarks_into(bufs, [cx] * len(sizes), dv["mus"], dv["sigma"], dv["rotx"], dv["roty"], dv["rotz"], 0, 0, -2.)
    for o, c in zip(outs, cat_ref):
        assert torch.equal(o[..., :cx].cpu(), c[..., :cx].detach()) and torch.equal(o[..., cx + K:].cpu(), c[..., cx + K:])
        assert_close_norm(o[..., cx:cx + K], c[..., cx:cx + K], FWD_RTOL, f"concat maps K={K} n={c.shape[1]}")
    sum((o * w.to(cuda_device)).sum() for o, w in zip(outs, ws)).backward()
    for a, b_ in zip(xd, xr):
        assert torch.equal(a.grad.cpu(), b_.grad)                       # pass-through channels: exact
    for k in ("mus", "sigma", "roty"):
        assert_close_norm(dv[k].grad, lv[k].grad, GRAD_RTOL, f"concat-direct grad {k} (K={K}, cx={cx})")
    # identical values to the dense path where that is the same kernel family
    dense = gg.get_landmarks(dv["mus"].detach(), dv["sigma"].detach(), dv["rotx"], dv["roty"].detach(), dv["rotz"], 0, 0, -2.,
                             sizes=sizes)
    for o, d_ in zip(outs, dense):
        assert_close_norm(o[..., cx:cx + K], d_, 1e-6, "concat-direct vs dense")


# ------------------------------------------------------------------------------------- several cameras per set
@pytest.mark.parametrize("B,K,V,sizes", [(3, 6, 3, (32, 64, 128, 256)), (3, 12, 3, (32, 64)), (2, 16, 2, (64,)), (5, 5, 4, (24,)),
                                         (1, 130, 2, (16,))])
def test_multi_view_render_vs_separate_oracle_calls(gg, cuda_device, B, K, V, sizes):
    """get_landmarks_views: V cameras on the same Gaussians in one projection + one splat launch (the reference's pose /
    pose + random yaw / canonical renders, mask/mask_gen_dynamic.py:707-709, at its own batch 3 and K = 6 / 12) against V
    separate oracle calls; d mus / d Sigma are the sums over the views, every view's angles get their own gradient."""
    inp = rp.synth_inputs(B, K, seed=900 + K)
    g = torch.Generator().manual_seed(B * K + V)
    rots = [((torch.rand(V, B, 1, generator=g) - 0.5) * a).float() for a in (30.0, 300.0, 30.0)]
    rots[1][V - 1] = 0.0                                                   # the canonical view: yaw 0
    ws = [[torch.randn(B, n, n, K, generator=g) for n in sizes] for _ in range(V)]
    wm = [torch.randn(B, sizes[-1], sizes[-1], 1, generator=g) for _ in range(V)]
    lv = {k: inp[k].clone().requires_grad_(True) for k in ("mus", "sigma")}
    lr = [r.clone().requires_grad_(True) for r in rots]
    total = 0
    ref = []
    for v in range(V):
        maps = rp.get_landmarks(lv["mus"], lv["sigma"], lr[0][v], lr[1][v], lr[2][v], 0, 0, -2., sizes=sizes)
        ref.append(maps)
        total = total + sum((m * w).sum() for m, w in zip(maps, ws[v])) + (maps[-1].sum(-1, keepdim=True) * wm[v]).sum()
    total.backward()
    dv = {k: inp[k].clone().to(cuda_device).requires_grad_(True) for k in ("mus", "sigma")}
    dr = [r.clone().to(cuda_device).requires_grad_(True) for r in rots]
    out = gg.get_landmarks_views(dv["mus"], dv["sigma"], dr[0], dr[1], dr[2], 0, 0, -2., sizes=sizes, with_mask=True)
    assert len(out) == V
    tot = 0
    for v in range(V):
        maps, mask = out[v]
        for m, r in zip(maps, ref[v]):
            assert_close_norm(m, r, FWD_RTOL, f"view {v} maps n={r.shape[1]}")
        assert_close_norm(mask, ref[v][-1].detach().sum(-1, keepdim=True), FWD_RTOL, f"view {v} silhouette")
        tot = tot + sum((m * w.to(cuda_device)).sum() for m, w in zip(maps, ws[v])) + (mask * wm[v].to(cuda_device)).sum()
    tot.backward()
    for k in ("mus", "sigma"):
        assert_close_norm(dv[k].grad, lv[k].grad, GRAD_RTOL, f"multi-view grad {k}")
    for i, nm in enumerate(("rotx", "roty", "rotz")):
        assert_close_norm(dr[i].grad, lr[i].grad, GRAD_RTOL, f"multi-view grad {nm}")
    # one view through the same entry equals the ordinary call bit for bit
    one = gg.get_landmarks_views(dv["mus"].detach(), dv["sigma"].detach(), dr[0][:1].detach(), dr[1][:1].detach(),
                                 dr[2][:1].detach(), 0, 0, -2., sizes=sizes)[0]
    plain = gg.get_landmarks(dv["mus"].detach(), dv["sigma"].detach(), dr[0][0].detach(), dr[1][0].detach(), dr[2][0].detach(),
                             0, 0, -2., sizes=sizes)
    for a, b_ in zip(one, plain):
        assert torch.equal(a, b_)


# ----------------------------------------------------------------------------------------------- red zones
def _zoned(shape, dtype, dev, fill):
    """A tensor of `shape` carved out of the middle of a larger allocation whose margins hold a sentinel."""
    n = int(np.prod(shape))
    pad = 4096                                                     # elements on either side
    pad -= pad % 64
    big = torch.full((n + 2 * pad,), fill, dtype=dtype, device=dev)
    return big, big[pad:pad + n].view(shape), pad


def _margins_intact(big, pad, fill):
    lo, hi = big[:pad], big[-pad:]
    if isinstance(fill, float) and math.isnan(fill):
        return bool(torch.isnan(lo).all() and torch.isnan(hi).all())
    return bool((lo == fill).all() and (hi == fill).all())


@pytest.mark.parametrize("B,K,sizes", [(3, 16, (32, 64, 128, 256)), (2, 6, (24, 40)), (1, 5, (17,)), (2, 64, (72,)), (2, 12, (48, 8))])
def test_no_entry_point_writes_outside_its_buffers(gg, cuda_device, restore_modes, B, K, sizes):
    """compute-sanitizer is closed on this pool (profiles/r2_sanitizer_unavailable.txt), so the out-of-bounds check is
    done with red zones: every output of every launch entry (projection, maps and masks of the forward in both
    layouts and both arithmetic modes, partials of the backward through the bulk-copy, quad, strip and fused kernels,
    loss partials, gradients) sits in the middle of a larger allocation filled with NaN, and the margins must still
    be NaN afterwards; inputs get the same treatment so that a read past their end would poison the results, which are
    checked to be finite."""
    from gaussigan_b200._lib import GG_MOMENTS, GG_PARAM_STRIDE, LAYOUT_NCHW, LAYOUT_NHWC, check, int_array, lib, ptr_array
    dev, nan = cuda_device, float("nan")
    st = torch.cuda.current_stream(dev).cuda_stream
    inp = rp.synth_inputs(B, K, seed=B * 31 + K)
    zoned_in = {}
    for k, shape, dt in (("mus", (B, K, 3), torch.float32), ("sigma", (B, K, 3, 3), torch.float64), ("rotx", (B,), torch.float32),
                         ("roty", (B,), torch.float32), ("rotz", (B,), torch.float32)):
        big, view, pad = _zoned(shape, dt, dev, nan)
        view.copy_(inp[k].reshape(shape).to(dev))
        zoned_in[k] = (big, view, pad)
    I = {k: v[1] for k, v in zoned_in.items()}
    outs = []

    def zone(shape, dtype=torch.float32):
        big, view, pad = _zoned(shape, dtype, dev, nan)
        outs.append((big, view, pad))
        return view

    params = zone((B, K, GG_PARAM_STRIDE))
    mu2d, sg2d = zone((B, K, 2), torch.float64), zone((B, K, 2, 2), torch.float64)
    check(lib.gg_project_fwd(I["mus"].data_ptr(), I["sigma"].data_ptr(), 1, I["rotx"].data_ptr(), I["roty"].data_ptr(),
                             I["rotz"].data_ptr(), 0., 0., -2., 1., B, K, mu2d.data_ptr(), sg2d.data_ptr(), params.data_ptr(),
                             dev.index, st), "project_fwd")
    sz = int_array(sizes)
    ns = len(sizes)
    for fast in (True, False):
        gg.set_fast_mask(fast)
        for layout in (LAYOUT_NHWC, LAYOUT_NCHW):
            maps = [zone((B, n, n, K) if layout == LAYOUT_NHWC else (B, K, n, n)) for n in sizes]
            masks = [zone((B, n, n)) for n in sizes]
            check(lib.gg_splat_fwd(params.data_ptr(), B, K, ns, sz, ptr_array([m.data_ptr() for m in maps]),
                                   ptr_array([m.data_ptr() for m in masks]), layout, dev.index, st), "splat_fwd")
            only_masks = [zone((B, n, n)) for n in sizes]
            check(lib.gg_splat_fwd(params.data_ptr(), B, K, ns, sz, ptr_array([None] * ns),
                                   ptr_array([m.data_ptr() for m in only_masks]), layout, dev.index, st), "splat_fwd(mask)")
            for m in maps + masks + only_masks:
                assert torch.isfinite(m).all()
            # backward: gradients of maps + masks, of maps only, of masks only
            for hm, hk in (((1,) * ns, (1,) * ns), ((1,) * ns, (0,) * ns), ((0,) * ns, (1,) * ns)):
                T = check(lib.gg_splat_bwd_tiles(K, ns, sz, int_array(hm), int_array(hk), layout), "tiles")
                part = zone((B, T, K, GG_MOMENTS))
                gm = [_zoned(tuple(m.shape), torch.float32, dev, nan) for m in maps]
                gk = [_zoned(tuple(m.shape), torch.float32, dev, nan) for m in masks]
                for _, v, _ in gm + gk:
                    v.normal_()
                check(lib.gg_splat_bwd(params.data_ptr(), B, K, ns, sz,
                                       ptr_array([g[1].data_ptr() if h else None for g, h in zip(gm, hm)]),
                                       ptr_array([g[1].data_ptr() if h else None for g, h in zip(gk, hk)]), layout,
                                       part.data_ptr(), T, dev.index, st), "splat_bwd")
                assert torch.isfinite(part).all(), (fast, layout, hm, hk)        # a read past an input would be NaN
                dmus, dsig, drot = zone((B, K, 3)), zone((B, K, 3, 3), torch.float64), zone((B, 3))
                check(lib.gg_project_bwd(I["mus"].data_ptr(), I["sigma"].data_ptr(), 1, I["rotx"].data_ptr(), I["roty"].data_ptr(),
                                         I["rotz"].data_ptr(), 0., 0., -2., 1., B, K, part.data_ptr(), T, None, None,
                                         dmus.data_ptr(), dsig.data_ptr(), drot.data_ptr(), dev.index, st), "project_bwd")
                assert torch.isfinite(dmus).all() and torch.isfinite(dsig).all() and torch.isfinite(drot).all()
        # fused silhouette loss (K <= 64), uint8 and float targets
        if K <= 64:
            n = sizes[-1]
            T = check(lib.gg_splat_bwd_tiles(K, 1, int_array([n]), int_array([0]), int_array([1]), LAYOUT_NHWC), "tiles")
            for dt, kind in ((torch.uint8, 0), (torch.float32, 1)):
                tb, tv, tp = _zoned((B, n, n), dt, dev, 255 if dt == torch.uint8 else nan)
                tv.copy_((torch.rand(B, n, n, device=dev) * (255 if dt == torch.uint8 else 1)).to(dt))
                part, lp, mo = zone((B, T, K, GG_MOMENTS)), zone((B * T,)), zone((B, n, n))
                check(lib.gg_silhouette_loss_fused(params.data_ptr(), tv.data_ptr(), kind, B, K, n, mo.data_ptr(), part.data_ptr(),
                                                   T, lp.data_ptr(), dev.index, st), "fused")
                assert torch.isfinite(part).all() and torch.isfinite(lp).all() and torch.isfinite(mo).all()
    # cycle loss and colourise
    n = sizes[0]
    Tc, nl = lib.gg_cycle_tiles(n), lib.gg_cycle_loss_count(B, K, n)
    qa, qb, lp = zone((B, Tc, K, GG_MOMENTS)), zone((B, Tc, K, GG_MOMENTS)), zone((nl,))
    check(lib.gg_cycle_l1_fused(params.data_ptr(), params.data_ptr(), B, K, n, qa.data_ptr(), qb.data_ptr(), Tc, lp.data_ptr(),
                                dev.index, st), "cycle")
    cols = torch.rand(K, 3, device=dev)
    rgb = zone((B, n, n, 3))
    check(lib.gg_colorize(params.data_ptr(), None, cols.data_ptr(), B, K, n, rgb.data_ptr(), 0, dev.index, st), "colorize")
    torch.cuda.synchronize()
    for big, view, pad in outs:
        assert _margins_intact(big, pad, nan), f"write outside a buffer of shape {tuple(view.shape)}"
    for k, (big, view, pad) in zoned_in.items():
        assert _margins_intact(big, pad, nan), f"input {k} was written"


# --------------------------------------------------------------------------------- lossless strip-packed targets
def _mask_images(B, n, kind, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "discs":                                   # two-level
        return ((rp.synth_target_masks(B, n, seed=seed).reshape(B, n, n) + 1.0) * 127.5).round().to(torch.uint8)
    if kind == "antialiased":                             # flat inside / outside, intermediate values on the edge
        ax = torch.linspace(-1, 1, n)
        yy, xx = torch.meshgrid(ax, ax, indexing="ij")
        c = torch.rand(B, 2, generator=g) - 0.5
        r = torch.rand(B, 1, 1, generator=g) * 0.4 + 0.2
        d = ((xx[None] - c[:, 0, None, None]) ** 2 + (yy[None] - c[:, 1, None, None]) ** 2).sqrt()
        return (255.0 * torch.clamp((r - d) / (3.0 / n) + 0.5, 0, 1)).round().to(torch.uint8)
    return torch.randint(0, 256, (B, n, n), generator=g, dtype=torch.uint8)       # noise: every strip raw


@pytest.mark.parametrize("n", [256, 100, 37, 8, 5, 520])
@pytest.mark.parametrize("kind", ["discs", "antialiased", "noise"])
def test_strip_packed_targets_equal_uint8_targets(gg, cuda_device, n, kind):
    """GG_TARGET_STRIPS (any 8-bit mask, two header bits per flat 8-pixel strip, 8 raw bytes otherwise) gives bit for bit
    the loss and gradients of the same image as raw uint8, through the planned fused step and the host-buffer step;
    the packed buffer is several times smaller for silhouette-like masks and never more than 7 % larger."""
    from gaussigan_b200.host import HostSilhouetteStep, encode_target_strips
    from gaussigan_b200.plan import SilhouetteLossStep
    B, K = 3, 16
    inp = rp.synth_inputs(B, K, seed=9 + n)
    raw = _mask_images(B, n, kind, seed=n)
    packed = encode_target_strips(raw, pin=False)
    if kind != "noise" and n >= 100:
        assert packed.numel() * 4 < raw.numel(), (packed.numel(), raw.numel())
    assert packed.numel() <= raw.numel() * 1.07 + 16 + 4 * B * n + 16 * B * n
    step = SilhouetteLossStep(B, K, n, cuda_device)
    step.bind(inp["mus"].reshape(B, K, 3).contiguous().to(cuda_device), inp["sigma"].to(cuda_device),
              *(inp[k].reshape(B).contiguous().to(cuda_device) for k in ("rotx", "roty", "rotz")))
    ref = [t.clone() for t in step.loss_backward(raw.to(cuda_device))] + [step.loss_partials.clone()]
    dev_packed = torch.empty((packed.numel() + 7) // 8, dtype=torch.int64, device=cuda_device).view(torch.uint8)
    dev_packed[:packed.numel()] = packed.to(cuda_device)
    out = [t.clone() for t in step.loss_backward(dev_packed, strips=True)] + [step.loss_partials.clone()]
    for a, b_ in zip(out, ref):
        assert torch.equal(a, b_)
    # the host-buffer step
    tgt_pm1 = (raw.float() / 127.5 - 1.0).reshape(B, n, n, 1)
    res = []
    for strips in (False, True):
        host = HostSilhouetteStep(B, K, n, cuda_device, depth=2, target_strips=strips)
        h = host.make_host_inputs(inp, tgt_pm1)
        r = host.submit(h).wait()
        res.append([r.loss_partials.clone(), r.dmus.clone(), r.dsigma.clone(), r.drot.clone()])
        if strips and kind != "noise" and n >= 100:
            assert host.h2d_bytes < 0.4 * (host.small_h2d_bytes + B * n * n)
    for a, b_ in zip(res[0], res[1]):
        assert torch.equal(a, b_)
