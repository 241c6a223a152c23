"""Parity of the CUDA path (through the public API and the C ABI) against the oracle and the golden vectors.

Tolerances (BASELINE.json north_star; norm-relative):
    forward  : max|out - ref| <= 1e-5 * max|ref|
    gradients: max|g - ref|   <= 1e-4 * max|ref|
Nothing here reads /root/reference (it does not exist on the GPU box).
"""
import os

import numpy as np
import pytest
import torch

from oracle import reference_port as rp
from util import FWD_RTOL, GRAD_RTOL, assert_close_norm, inputs_of, load, max_rel, sub, weights_like

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gg(cuda_device):
    import gaussigan_b200
    return gaussigan_b200


def _dev(inp, dev):
    return {k: v.to(dev) for k, v in inp.items()}


def _leaves(inp, dev=None):
    return {k: (v.to(dev) if dev is not None else v.clone()).requires_grad_(True) for k, v in inp.items()}


# ------------------------------------------------------------------------------ golden vectors
@pytest.mark.parametrize("name", ["train_k6", "train_k12_general", "train_k16"])
@pytest.mark.parametrize("layout", ["nhwc", "nchw"])
def test_get_landmarks_vs_golden(gg, cuda_device, name, layout):
    g = load(name)
    lv = _leaves(inputs_of(g), cuda_device)
    t = [float(v) for v in g["t"]]
    maps = gg.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], *t, float(g["focal"]),
                            layout=layout)
    if layout == "nchw":
        maps = [m.permute(0, 2, 3, 1) for m in maps]
    assert [m.shape[1] for m in maps] == [32, 64, 128, 256] and all(m.dtype == torch.float32 for m in maps)
    for m in maps:
        n = m.shape[1]
        assert_close_norm(sub(m), g[f"maps{n}"], FWD_RTOL, f"{name} maps{n}")
        assert_close_norm(m.double().sum((1, 2)), g[f"maps{n}_sum"], FWD_RTOL, f"{name} sum{n}")
    w = weights_like([m.shape for m in maps], g["wseed"])
    sum((m * q.to(cuda_device)).sum() for m, q in zip(maps, w)).backward()
    for k, v in lv.items():
        assert v.grad.dtype == v.dtype and v.grad.shape == v.shape
        assert_close_norm(v.grad, g[f"grad_{k}"], GRAD_RTOL, f"{name} grad_{k}")


def test_infer_flavour_vs_golden(gg, cuda_device):
    g = load("infer_k6")
    d = _dev(inputs_of(g), cuda_device)
    mu2d, s2d, maps = gg.get_landmarks_infer(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0., 0., -2.)
    assert mu2d.dtype == s2d.dtype == maps.dtype == torch.float32 and tuple(maps.shape) == (2, 256, 256, 6)
    np.testing.assert_array_equal(mu2d.cpu().numpy(), g["mu2d"])          # f64 projection -> same f32 rounding
    np.testing.assert_array_equal(s2d.cpu().numpy(), g["sigma2d"])
    assert_close_norm(sub(maps), g["maps256"], FWD_RTOL, "infer maps")


def test_get_gaussian_maps_2d_vs_golden(gg, cuda_device):
    g = load("raster_k8")
    n = int(g["n"])
    mu = torch.from_numpy(g["mu2d"]).to(cuda_device).requires_grad_(True)
    sg = torch.from_numpy(g["sigma2d"]).to(cuda_device).requires_grad_(True)
    m = gg.get_gaussian_maps_2d(mu, sg, [n, n])
    assert_close_norm(m, g["maps"], FWD_RTOL, "raster maps")
    w = weights_like([m.shape], g["wseed"])[0].to(cuda_device)
    (m * w).sum().backward()
    assert mu.grad.dtype == torch.float64
    assert_close_norm(mu.grad, g["grad_mu2d"], GRAD_RTOL, "grad_mu2d")
    assert_close_norm(sg.grad, g["grad_sigma2d"], GRAD_RTOL, "grad_sigma2d")
    with pytest.raises(ValueError):
        gg.get_gaussian_maps_2d(mu, sg, [32, 48])


def test_randomised_get_gaussian_maps_2d(gg, cuda_device):
    """Seeded sweep of the raster-only boundary (the frozen-graph inputs mu2d / sigma2d, mask/infer_masks.py:394-395):
    K 1..24, any n, float32 and float64 inputs, positive-definite 2x2 covariances from a pixel to the frame size;
    maps and the gradients w.r.t. mu2d and sigma2d against the oracle."""
    rng = torch.Generator().manual_seed(31)
    for case in range(int(os.environ.get("GG_SWEEP_CASES", "100"))):
        B = int(torch.randint(1, 4, (1,), generator=rng))
        K = int(torch.randint(1, 25, (1,), generator=rng))
        n = int(torch.randint(2, 150, (1,), generator=rng))
        dt = torch.float64 if bool(torch.randint(0, 2, (1,), generator=rng)) else torch.float32
        mu = ((torch.rand(B, K, 2, generator=rng, dtype=torch.float64) - 0.5) * 2.4)
        ang = torch.rand(B, K, generator=rng, dtype=torch.float64) * 3.14159
        lo, hi = torch.log(torch.tensor(4.0 / n ** 2, dtype=torch.float64)), torch.log(torch.tensor(0.5, dtype=torch.float64))
        var = torch.exp(torch.rand(B, K, 2, generator=rng, dtype=torch.float64) * (hi - lo) + lo)    # sigma >= 1 pixel
        c, s_ = torch.cos(ang), torch.sin(ang)
        R = torch.stack([torch.stack([c, -s_], -1), torch.stack([s_, c], -1)], -2)
        sg = (R * var[..., None, :]) @ R.transpose(-1, -2)
        mu, sg = mu.to(dt), sg.to(dt)
        lm, ls = mu.double().clone().requires_grad_(True), sg.double().clone().requires_grad_(True)
        ref = rp.get_gaussian_maps_2d(lm, ls, [n, n])
        w = torch.randn(B, n, n, K, generator=rng)
        (ref * w).sum().backward()
        dm, ds = mu.to(cuda_device).requires_grad_(True), sg.to(cuda_device).requires_grad_(True)
        out = gg.get_gaussian_maps_2d(dm, ds, [n, n])
        what = f"case {case} (B={B}, K={K}, n={n}, {dt})"
        assert_close_norm(out, ref, FWD_RTOL, what + " maps")
        (out * w.to(cuda_device)).sum().backward()
        assert dm.grad.dtype == dt and ds.grad.dtype == dt
        assert_close_norm(dm.grad, lm.grad, GRAD_RTOL, what + " grad mu2d")
        assert_close_norm(ds.grad, ls.grad, GRAD_RTOL, what + " grad sigma2d")


# ------------------------------------------------------------------- BASELINE.json configurations
def test_config1_forward_b16_k8_128(gg, cuda_device):
    inp = rp.synth_inputs(16, 8, seed=56)
    d = _dev(inp, cuda_device)
    ref = rp.get_landmarks(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0, 0, -2., sizes=(128,))[0]
    out = gg.get_landmarks(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., sizes=(128,))[0]
    assert_close_norm(out, ref, FWD_RTOL, "C1 maps")
    mask = gg.render_silhouette(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., size=128)
    assert_close_norm(mask, ref.sum(-1), FWD_RTOL, "C1 mask")


def test_config2_forward_backward_b64_k16_256_full_size(gg, cuda_device):
    """The headline configuration at full size: mask renderer fwd+bwd, gradient check vs the oracle."""
    B, K, n = 64, 16, 256
    inp = rp.synth_inputs(B, K, seed=56)
    tgt = rp.synth_target_masks(B, n)
    mask_ref, loss_ref, dmu_ref, dsig_ref, drot_ref = rp.render_loss_and_grads(inp, tgt, n)
    d = _dev(inp, cuda_device)
    mus, sig, roty = (d[k].clone().requires_grad_(True) for k in ("mus", "sigma", "roty"))
    mask = gg.render_silhouette(mus, sig, d["rotx"], roty, d["rotz"], 0, 0, -2., size=n)
    loss = ((tgt.to(cuda_device)[..., 0] + 1) * 0.5 - mask).abs().mean()
    loss.backward()
    assert_close_norm(mask, mask_ref, FWD_RTOL, "C2 mask")
    assert abs(loss.item() - loss_ref.item()) <= 1e-6 * abs(loss_ref.item())
    assert_close_norm(mus.grad, dmu_ref, GRAD_RTOL, "C2 dmus")
    assert_close_norm(sig.grad, dsig_ref, GRAD_RTOL, "C2 dsigma (raw, as the reference's autodiff)")
    assert_close_norm(rp.sym(sig.grad), rp.sym(dsig_ref), GRAD_RTOL, "C2 sym(dsigma)")
    assert_close_norm(roty.grad, drot_ref, GRAD_RTOL, "C2 droty")


def test_config3_deformation_then_render(gg, cuda_device):
    """Dynamic-object step: per-frame deformation (reference tail, here torch ops) feeding the renderer;
    gradients flow back through the renderer into the canonical Gaussians and the deformation heads."""
    B, K, n = 8, 16, 256
    g = torch.Generator().manual_seed(56)
    th = lambda *s: torch.tanh(torch.randn(*s, generator=g))
    cann = rp.synth_inputs(1, K, seed=56)
    heads = [th(B, K * 3), th(B, K * 3), th(B, K), th(B, K), th(B, K), th(B, 1)]
    w = torch.randn(B, n, n, generator=g)

    def run(mod, dev, render):
        cm = cann["mus"].detach().clone().to(dev).requires_grad_(True)
        cs = cann["sigma"].detach().clone().to(dev).requires_grad_(True)
        hs = [h.detach().clone().to(dev).requires_grad_(True) for h in heads]
        mus, sig, theta = mod.deform(*hs, cm, cs)
        m = render(mus, sig, torch.zeros_like(theta), theta, torch.zeros_like(theta))
        (m * w.to(dev)).sum().backward()
        return m.detach(), [cm.grad, cs.grad] + [h.grad for h in hs]

    m_ref, g_ref = run(rp, torch.device("cpu"),
                       lambda mu, s, rx, ry, rz: rp.get_landmarks(mu, s, rx, ry, rz, 0, 0, -2., sizes=(n,))[0].sum(-1))
    m_out, g_out = run(gg, cuda_device, lambda mu, s, rx, ry, rz: gg.render_silhouette(mu, s, rx, ry, rz, 0, 0, -2., size=n))
    # the library's own deformation and Jacobi SVD kernels on the GPU side (float32 trig pinned to the oracle's
    # correctly rounded definition): the path's own tolerances, no slack
    assert_close_norm(m_out, m_ref, FWD_RTOL, "C3 mask")
    for a, b, nm in zip(g_out, g_ref, ["cannmu", "cannsigma", "mu_t", "scale", "rx", "ry", "rz", "theta"]):
        assert_close_norm(a, b, GRAD_RTOL, "C3 grad " + nm)


def test_config4_maps_pyramid_k16(gg, cuda_device):
    """rgb path: K=16 per-Gaussian maps at the four scales (forward only in the reference, rgb_gen_dynamic.py:956)."""
    B, K = 32, 16
    inp = rp.synth_inputs(B, K, seed=4, angle=60.0)
    d = _dev(inp, cuda_device)
    ref = rp.get_landmarks(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0, 0, -2.)
    out, mask = gg.get_landmarks_with_mask(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2.)
    for o, r in zip(out, ref):
        assert_close_norm(o, r, FWD_RTOL, f"C4 maps{r.shape[1]}")
    assert tuple(mask.shape) == (B, 256, 256, 1)
    assert_close_norm(mask, rp.sum_composite(ref[-1]), FWD_RTOL, "C4 fused sum composite")


def test_config5_shape_k64_512_properties_and_oracle(gg, cuda_device):
    """Throughput-sweep shape (K=64, 512x512) at a batch the oracle finishes in seconds, plus
    size-independent properties."""
    B, K, n = 2, 64, 512
    inp = rp.synth_inputs(B, K, seed=5)
    d = _dev(inp, cuda_device)
    lv = _leaves(inp)
    ref = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0, 0, -2., sizes=(n,))[0]
    gen = torch.Generator().manual_seed(1)
    gm = torch.randn(B, n, n, generator=gen)
    (ref.sum(-1) * gm).sum().backward()
    mus, sig, roty = (d[k].clone().requires_grad_(True) for k in ("mus", "sigma", "roty"))
    mask = gg.render_silhouette(mus, sig, d["rotx"], roty, d["rotz"], 0, 0, -2., size=n)
    mask.backward(gm.to(cuda_device))
    assert_close_norm(mask, ref.sum(-1), FWD_RTOL, "C5 mask")
    assert_close_norm(mus.grad, lv["mus"].grad, GRAD_RTOL, "C5 dmus")
    assert_close_norm(sig.grad, lv["sigma"].grad, GRAD_RTOL, "C5 dsigma")
    assert_close_norm(roty.grad, lv["roty"].grad, GRAD_RTOL, "C5 droty")
    # property: K-channel maps (quad kernel, K=64 -> 16 lanes per pixel) sum to the mask-mode kernel's output
    maps = gg.get_landmarks(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., sizes=(n,))[0]
    assert_close_norm(maps.sum(-1), mask, 2e-6, "maps.sum == mask")
    assert float(maps.max()) <= 1.0 + 1e-6 and float(maps.min()) >= 0.0


def test_backward_is_linear_and_deterministic(gg, cuda_device):
    B, K, n = 16, 16, 256
    d = _dev(rp.synth_inputs(B, K, seed=8), cuda_device)
    gen = torch.Generator().manual_seed(2)
    g1 = torch.randn(B, n, n, generator=gen).to(cuda_device)
    g2 = torch.randn(B, n, n, generator=gen).to(cuda_device)

    def grads(gm):
        mus, sig, roty = (d[k].clone().requires_grad_(True) for k in ("mus", "sigma", "roty"))
        gg.render_silhouette(mus, sig, d["rotx"], roty, d["rotz"], 0, 0, -2., size=n).backward(gm)
        return mus.grad, sig.grad, roty.grad

    a, b, c, a2 = grads(g1), grads(g2), grads(2.0 * g1 - 3.0 * g2), grads(g1)
    for x, y, z, x2 in zip(a, b, c, a2):
        assert torch.equal(x, x2)                                  # no atomics: bit-reproducible
        assert_close_norm(z, 2.0 * x.double() - 3.0 * y.double(), 1e-5, "linearity of the adjoint")


# ----------------------------------------------------------------------------------- edge cases
@pytest.mark.parametrize("B,K,sizes", [
    (1, 1, (1,)), (1, 1, (2,)), (2, 3, (7, 33)), (2, 5, (100,)), (1, 6, (12, 20, 36, 44)),
    (2, 12, (64,)), (2, 80, (40,)), (1, 256, (24,)), (1, 2, (1000,)), (1, 2, (2100,)),
])
@pytest.mark.parametrize("layout", ["nhwc", "nchw"])
def test_ragged_shapes_forward_backward(gg, cuda_device, B, K, sizes, layout):
    inp = rp.synth_inputs(B, K, seed=B * 1000 + K, angle=30.0)
    lv = _leaves(inp)
    ref = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0, 0, -2., sizes=sizes)
    w = weights_like([r.shape for r in ref], 9)
    sum((r * q).sum() for r, q in zip(ref, w)).backward()
    dv = _leaves(inp, cuda_device)
    out = gg.get_landmarks(dv["mus"], dv["sigma"], dv["rotx"], dv["roty"], dv["rotz"], 0, 0, -2., sizes=sizes, layout=layout)
    if layout == "nchw":
        out = [o.permute(0, 2, 3, 1) for o in out]
    for o, r in zip(out, ref):
        assert o.shape == r.shape
        assert_close_norm(o, r, FWD_RTOL, f"maps n={r.shape[1]}")
    sum((o * q.to(cuda_device)).sum() for o, q in zip(out, w)).backward()
    for k in ("mus", "sigma", "rotx", "roty", "rotz"):
        assert_close_norm(dv[k].grad, lv[k].grad, GRAD_RTOL, f"grad {k}")
    # mask-only path on the last scale
    m = gg.render_silhouette(dv["mus"].detach(), dv["sigma"].detach(), dv["rotx"].detach(), dv["roty"].detach(),
                             dv["rotz"].detach(), 0, 0, -2., size=sizes[-1])
    assert_close_norm(m, ref[-1].detach().sum(-1), FWD_RTOL, "mask")


def test_empty_batch(gg, cuda_device):
    z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=cuda_device)
    maps = gg.get_landmarks(z(0, 4, 1, 3), z(0, 4, 3, 3, dt=torch.float64), z(0, 1), z(0, 1), z(0, 1), 0, 0, -2.)
    assert [tuple(m.shape) for m in maps] == [(0, 32, 32, 4), (0, 64, 64, 4), (0, 128, 128, 4), (0, 256, 256, 4)]
    assert tuple(gg.render_silhouette(z(0, 4, 1, 3), z(0, 4, 3, 3), z(0, 1), z(0, 1), z(0, 1), 0, 0, -2.).shape) == (0, 256, 256)


def test_float32_sigma_and_flat_shapes(gg, cuda_device):
    inp = rp.synth_inputs(3, 8, seed=2)
    s32 = inp["sigma"].float()
    ref = rp.get_landmarks(inp["mus"], s32.double(), inp["rotx"], inp["roty"], inp["rotz"], 0, 0, -2., sizes=(64,))[0]
    d = _dev(inp, cuda_device)
    sig = s32.to(cuda_device).requires_grad_(True)
    out = gg.get_landmarks(d["mus"].reshape(3, 8, 3), sig, d["rotx"].reshape(3), d["roty"].reshape(3),
                           d["rotz"].reshape(3), 0, 0, -2., sizes=(64,), K=8)[0]
    assert_close_norm(out, ref, FWD_RTOL, "f32 sigma")
    out.sum().backward()
    assert sig.grad.dtype == torch.float32 and sig.grad.shape == sig.shape
    with pytest.raises(ValueError):
        gg.get_landmarks(d["mus"], sig, d["rotx"], d["roty"], d["rotz"], 0, 0, -2., K=7)


def test_offscreen_and_tiny_gaussians_stay_finite(gg, cuda_device):
    """Gaussians far outside the frame or much smaller than a pixel: zeros / tiny values, never NaN."""
    B, K, n = 2, 4, 64
    mus = torch.tensor([[[3.0, 0, 0], [0, -3.0, 0], [0.0, 0, 0], [0.3, 0.3, 0]]] * B).view(B, K, 1, 3)
    sig = torch.eye(3, dtype=torch.float64).view(1, 1, 3, 3).repeat(B, K, 1, 1) * 0.01
    sig[:, 2] *= 1e-4
    z = torch.zeros(B, 1)
    ref = rp.get_landmarks(mus, sig, z, z, z, 0, 0, -2., sizes=(n,))[0]
    dv = [t.to(cuda_device).requires_grad_(True) for t in (mus, sig)]
    zz = z.to(cuda_device)
    out = gg.get_landmarks(dv[0], dv[1], zz, zz, zz, 0, 0, -2., sizes=(n,))[0]
    assert torch.isfinite(out).all()
    assert_close_norm(out, ref, FWD_RTOL, "offscreen/tiny")
    out.sum().backward()
    assert torch.isfinite(dv[0].grad).all() and torch.isfinite(dv[1].grad).all()


# --------------------------------------------------------------------- planned step / C ABI direct
def test_planned_step_equals_autograd_and_graph_replay(gg, cuda_device):
    from gaussigan_b200.plan import SilhouetteStep
    B, K, n = 8, 16, 128
    d = _dev(rp.synth_inputs(B, K, seed=12), cuda_device)
    gm = torch.randn(B, n, n, device=cuda_device)
    mus = d["mus"].reshape(B, K, 3).contiguous()
    rx, ry, rz = (d[k].reshape(B).contiguous() for k in ("rotx", "roty", "rotz"))
    step = SilhouetteStep(B, K, n, cuda_device)
    mask = step.forward(mus, d["sigma"], rx, ry, rz).clone()
    dmus, dsig, drot = (t.clone() for t in step.backward(gm))
    lm, ls, lr = (t.clone().requires_grad_(True) for t in (d["mus"], d["sigma"], d["roty"]))
    m2 = gg.render_silhouette(lm, ls, d["rotx"], lr, d["rotz"], 0, 0, -2., size=n)
    m2.backward(gm)
    assert torch.equal(mask, m2) and torch.equal(dmus.view_as(lm.grad), lm.grad)
    assert torch.equal(dsig, ls.grad) and torch.equal(drot[:, 1], lr.grad.view(-1))
    step.capture(gm)
    step.mask.zero_(); step.dmus.zero_(); step.dsigma.zero_()
    step.graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(step.mask, mask) and torch.equal(step.dmus, dmus) and torch.equal(step.dsigma, dsig)


def test_projection_outputs_match_oracle_to_float64_roundoff(gg, cuda_device):
    inp = rp.synth_inputs(16, 16, seed=77, angle=90.0)
    inp["rotx"] = torch.randn(16, 1) * 20
    inp["rotz"] = torch.randn(16, 1) * 20
    d = _dev(inp, cuda_device)
    mu2d, s2d = gg.project(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0.3, -0.1, -2.4, 1.7)
    rmu, rs = rp.project(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0.3, -0.1, -2.4, 1.7)
    assert mu2d.dtype == torch.float64
    assert (mu2d.cpu() - rmu).abs().max() < 1e-12 and max_rel(s2d, rs) < 1e-12


@pytest.mark.parametrize("B,K,n", [(5, 16, 64), (3, 6, 40), (2, 21, 32), (4, 1, 24), (2, 12, 96)])
def test_linearised_projection_adjoint_equals_generic_kernel(gg, cuda_device, B, K, n):
    """project_bwd_lin_kernel (one camera per set, 6 K <= 128: Jacobian block per Gaussian ahead of the dependency wait,
    dot products after it) against project_bwd_kernel (GG_PROJ_BWD_LIN=0, read per call) on the same partial sums, all
    three Euler angles, with and without upstream gradients of mu2d / sigma2d on top: float64 round-off apart, and
    both within the gradient tolerance of the oracle."""
    inp = rp.synth_inputs(B, K, seed=900 + K, angle=60.0)
    inp["rotx"] = torch.randn(B, 1, generator=torch.Generator().manual_seed(1)) * 25
    inp["rotz"] = torch.randn(B, 1, generator=torch.Generator().manual_seed(2)) * 25
    w = weights_like([(B, n, n, 1)], 41)[0]
    lv = _leaves(inp)
    ref = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0.1, -0.2, -2.2, 1.3, sizes=(n,))[0]
    (ref.sum(-1, keepdim=True) * w).sum().backward()
    saved = os.environ.get("GG_PROJ_BWD_LIN")
    grads = {}
    try:
        for mode in ("1", "0"):
            os.environ["GG_PROJ_BWD_LIN"] = mode
            dv = _leaves(inp, cuda_device)
            m = gg.render_silhouette(dv["mus"], dv["sigma"], dv["rotx"], dv["roty"], dv["rotz"], 0.1, -0.2, -2.2, 1.3, size=n)
            (m.reshape(B, n, n, 1) * w.to(cuda_device)).sum().backward()
            grads[mode] = {k: dv[k].grad.clone() for k in ("mus", "sigma", "rotx", "roty", "rotz")}
            # the frozen-graph boundary tensors carry their own upstream gradients through the same kernel
            dv2 = _leaves(inp, cuda_device)
            mu2d, s2d = gg.project(dv2["mus"], dv2["sigma"], dv2["rotx"], dv2["roty"], dv2["rotz"], 0.1, -0.2, -2.2, 1.3)
            ((mu2d ** 2).sum() + (s2d * torch.arange(4, device=cuda_device, dtype=s2d.dtype).reshape(2, 2)).sum()).backward()
            grads[mode + "p"] = {k: dv2[k].grad.clone() for k in ("mus", "sigma", "roty")}
    finally:
        if saved is None:
            os.environ.pop("GG_PROJ_BWD_LIN", None)
        else:
            os.environ["GG_PROJ_BWD_LIN"] = saved
    for k in grads["1"]:
        assert_close_norm(grads["1"][k], lv[k].grad, GRAD_RTOL, f"K={K} grad {k} (linearised)")
        assert_close_norm(grads["0"][k], lv[k].grad, GRAD_RTOL, f"K={K} grad {k} (generic)")
        tol = 1e-6 if grads["1"][k].dtype == torch.float32 else 1e-9        # float32 outputs are rounded once
        assert_close_norm(grads["1"][k], grads["0"][k], tol, f"K={K} grad {k}: linearised vs generic")
    for k in grads["1p"]:
        tol = 1e-6 if grads["1p"][k].dtype == torch.float32 else 1e-9
        assert_close_norm(grads["1p"][k], grads["0p"][k], tol, f"K={K} projection-only grad {k}: linearised vs generic")


# ------------------------------------------------------------------ fused consumers / host-buffer step
@pytest.mark.parametrize("B,K,n,dtype", [(6, 16, 256, torch.uint8), (3, 6, 64, torch.float32), (2, 5, 100, torch.uint8),
                                         (2, 64, 40, torch.float32)])
def test_fused_silhouette_loss_step_vs_oracle(gg, cuda_device, B, K, n, dtype):
    from gaussigan_b200.plan import SilhouetteLossStep
    inp = rp.synth_inputs(B, K, seed=31 + K)
    tgt = rp.synth_target_masks(B, n, seed=5)                  # {-1,+1}
    if dtype == torch.uint8:
        raw = ((tgt + 1.0) * 127.5).round().clamp(0, 255)
        raw[:, ::7, ::5] = 93.0                                 # grey levels too, like a resized mask
        tgt = raw / 127.5 - 1.0                                 # what the reference sees (:672)
        dev_t = raw.reshape(B, n, n).to(torch.uint8).to(cuda_device)
    else:
        tgt = tgt * 0.8 + 0.1 * torch.randn(tgt.shape, generator=torch.Generator().manual_seed(1))
        dev_t = tgt.reshape(B, n, n).float().contiguous().to(cuda_device)
    mask_ref, loss_ref, dmu_ref, dsig_ref, drot_ref = rp.render_loss_and_grads(inp, tgt, n)
    d = _dev(inp, cuda_device)
    step = SilhouetteLossStep(B, K, n, cuda_device, want_mask=True)
    step.bind(d["mus"].reshape(B, K, 3).contiguous(), d["sigma"], *(d[k].reshape(B).contiguous() for k in ("rotx", "roty", "rotz")))
    dmus, dsig, drot = step.loss_backward(dev_t)
    assert abs(step.loss() - float(loss_ref)) <= 2e-6 * abs(float(loss_ref))
    assert_close_norm(step.mask, mask_ref, FWD_RTOL, "fused mask")
    assert_close_norm(dmus, dmu_ref.reshape(B, K, 3), GRAD_RTOL, "fused dmus")
    assert_close_norm(dsig, dsig_ref, GRAD_RTOL, "fused dsigma")
    assert_close_norm(drot[:, 1], drot_ref.reshape(B), GRAD_RTOL, "fused droty")
    # same numbers as the three-kernel path (splat fwd -> loss kernel -> splat bwd), bit for bit
    plain = gg.render_silhouette(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., size=n)
    assert torch.equal(plain, step.mask)


def test_host_buffer_step_matches_oracle(gg, cuda_device):
    from gaussigan_b200.host import HostSilhouetteStep
    B, K, n = 8, 16, 128
    inp = rp.synth_inputs(B, K, seed=77)
    tgt = rp.synth_target_masks(B, n, seed=3)
    mask_ref, loss_ref, dmu_ref, dsig_ref, drot_ref = rp.render_loss_and_grads(inp, tgt, n)
    host = HostSilhouetteStep(B, K, n, cuda_device, depth=2, want_mask=True)
    h = host.make_host_inputs(inp, tgt)
    results = [host.submit(h) for _ in range(3)]               # reuses slots: exercises the back-pressure path
    res = results[-1].wait()
    assert abs(res.loss - float(loss_ref)) <= 2e-6 * abs(float(loss_ref))
    assert_close_norm(res.mask, mask_ref, FWD_RTOL, "host mask")
    assert_close_norm(res.dmus, dmu_ref.reshape(B, K, 3), GRAD_RTOL, "host dmus")
    assert_close_norm(res.dsigma, dsig_ref, GRAD_RTOL, "host dsigma")
    assert_close_norm(res.drot[:, 1], drot_ref.reshape(B), GRAD_RTOL, "host droty")
    assert host.h2d_bytes == B * K * 3 * 4 + B * K * 9 * 8 + 3 * B * 4 + B * n * n and host.LAUNCHES_PER_STEP == 3


def test_randomised_host_step_equals_device_step(gg, cuda_device):
    """Seeded sweep: the host-buffer step (pinned host arrays in, gradients out; every GG_ZEROCOPY mode -- gradients
    written in place (default), inputs read in place, both, neither; uint8 / float32 / bit-packed targets; K <= 64
    fused, K > 64 five launches) returns exactly what the same kernels return when driven from device tensors."""
    from gaussigan_b200.host import HostSilhouetteStep
    from gaussigan_b200.plan import SilhouetteLossStep
    rng = torch.Generator().manual_seed(9)
    saved = os.environ.get("GG_ZEROCOPY")
    try:
        _host_step_sweep(gg, cuda_device, rng, HostSilhouetteStep, SilhouetteLossStep)
    finally:
        if saved is None:
            os.environ.pop("GG_ZEROCOPY", None)
        else:
            os.environ["GG_ZEROCOPY"] = saved


def _host_step_sweep(gg, cuda_device, rng, HostSilhouetteStep, SilhouetteLossStep):
    for case in range(int(os.environ.get("GG_SWEEP_CASES", "100"))):
        os.environ["GG_ZEROCOPY"] = ("out", "in", "1", "0")[case % 4]      # read by the library on every call
        B = int(torch.randint(1, 6, (1,), generator=rng))
        K = int(torch.randint(1, 65, (1,), generator=rng))
        n = int(torch.randint(1, 130, (1,), generator=rng))
        kind = int(torch.randint(0, 3, (1,), generator=rng))          # 0: uint8, 1: float32, 2: bits
        inp = rp.synth_inputs(B, K, seed=2000 + case)
        tgt = rp.synth_target_masks(B, n, seed=case)                  # {-1,+1}: exact in all three formats
        host = HostSilhouetteStep(B, K, n, cuda_device, depth=2, want_mask=True,
                                  target_dtype=torch.float32 if kind == 1 else torch.uint8, target_bits=kind == 2)
        h = host.make_host_inputs(inp, tgt)
        res = host.submit(h).wait()
        d = _dev(inp, cuda_device)
        step = SilhouetteLossStep(B, K, n, cuda_device, want_mask=True)
        step.bind(d["mus"].reshape(B, K, 3).contiguous(), d["sigma"], *(d[k].reshape(B).contiguous() for k in ("rotx", "roty", "rotz")))
        dm, ds, dr = step.loss_backward(h.target.to(cuda_device), bits=kind == 2)
        torch.cuda.synchronize()
        what = f"case {case} (B={B}, K={K}, n={n}, kind={kind})"
        assert torch.equal(res.mask.cpu(), step.mask.cpu()), what
        assert torch.equal(res.dmus.cpu(), dm.cpu()) and torch.equal(res.dsigma.cpu(), ds.cpu()), what
        assert torch.equal(res.drot.cpu(), dr.cpu()), what
        assert abs(res.loss - step.loss()) <= 1e-6 * max(abs(step.loss()), 1e-30), what


def test_unfused_loss_kernel_matches_fused(gg, cuda_device):
    """gg_mask_l1_loss_bwd (used when K > 64) against the oracle's loss and Abs gradient."""
    from gaussigan_b200._lib import lib, check
    B, n = 4, 64
    g = torch.Generator().manual_seed(0)
    mask = (torch.rand(B, n, n, generator=g) * 2).to(cuda_device)
    raw = torch.randint(0, 256, (B, n, n), generator=g, dtype=torch.uint8)
    gm = torch.empty_like(mask)
    lp = torch.empty(296, device=cuda_device)
    check(lib.gg_mask_l1_loss_bwd(mask.data_ptr(), raw.to(cuda_device).data_ptr(), 0, B, n, gm.data_ptr(), lp.data_ptr(),
                                  0, torch.cuda.current_stream().cuda_stream), "loss")
    m = mask.cpu().clone().requires_grad_(True)
    loss = rp.mask_recon_loss_bis((raw.float() / 127.5 - 1.0).unsqueeze(-1), m.unsqueeze(-1).expand(B, n, n, 1))
    loss.backward()
    assert abs(float(lp.double().sum()) / (B * n * n) - float(loss)) < 1e-6
    assert torch.equal(gm.cpu(), m.grad)


def test_randomised_fused_loss_step_equals_three_kernel_path(gg, cuda_device):
    """Seeded sweep over (B, K <= 64, n) and the three target formats: the fused silhouette + L1 + backward kernel
    gives the same silhouette and the same gradients, bit for bit, as splat forward -> loss kernel -> splat backward
    (same arithmetic, same reduction order), and the same loss up to the order of the per-CTA partial sums."""
    from gaussigan_b200._lib import lib, check
    from gaussigan_b200.plan import SilhouetteLossStep, SilhouetteStep
    rng = torch.Generator().manual_seed(5)
    for case in range(int(os.environ.get("GG_SWEEP_CASES", "100"))):
        B = int(torch.randint(1, 6, (1,), generator=rng))
        K = int(torch.randint(1, 65, (1,), generator=rng))
        n = int(torch.randint(1, 150, (1,), generator=rng))
        kind = int(torch.randint(0, 3, (1,), generator=rng))          # 0: uint8, 1: float32, 2: bits
        inp = rp.synth_inputs(B, K, seed=1000 + case)
        d = _dev(inp, cuda_device)
        args = (d["mus"].reshape(B, K, 3).contiguous(), d["sigma"], *(d[k].reshape(B).contiguous() for k in ("rotx", "roty", "rotz")))
        raw = torch.randint(0, 256, (B, n, n), generator=rng, dtype=torch.uint8)
        if kind == 0:
            tgt = raw.to(cuda_device)
        elif kind == 1:
            tgt = (raw.float() / 127.5 - 1.0).to(cuda_device)
        else:
            raw = (raw > 127).to(torch.uint8) * 255
            tgt = torch.from_numpy(np.packbits((raw > 0).numpy(), axis=-1, bitorder="little")).to(cuda_device)
        fused = SilhouetteLossStep(B, K, n, cuda_device, want_mask=True)
        fused.bind(*args)
        fdm, fds, fdr = (t.clone() for t in fused.loss_backward(tgt, bits=kind == 2))
        plain = SilhouetteStep(B, K, n, cuda_device)
        plain.bind(*args)
        mask = plain.forward()
        gm = torch.empty_like(mask)
        lp = torch.empty(296, device=cuda_device)
        t3 = tgt if kind != 2 else raw.to(cuda_device)                  # the loss kernel takes uint8 / float32
        check(lib.gg_mask_l1_loss_bwd(mask.data_ptr(), t3.data_ptr(), 1 if kind == 1 else 0, B, n, gm.data_ptr(),
                                      lp.data_ptr(), cuda_device.index or 0,
                                      torch.cuda.current_stream().cuda_stream), "loss")
        pdm, pds, pdr = plain.backward(gm)
        what = f"case {case} (B={B}, K={K}, n={n}, kind={kind})"
        assert torch.equal(fused.mask, mask), what
        assert torch.equal(fdm, pdm) and torch.equal(fds, pds) and torch.equal(fdr, pdr), what
        loss3 = float(lp.double().sum()) / (B * n * n)
        assert abs(fused.loss() - loss3) <= 1e-6 * max(abs(loss3), 1e-30), what


# ---------------------------------------------------------------- consumers
def test_mask_recon_loss_bis_autograd(gg, cuda_device):
    B, K, n = 5, 12, 128
    inp = rp.synth_inputs(B, K, seed=19, angle=20.0)
    tgt = rp.synth_target_masks(B, n, seed=8)
    lv = _leaves(inp)
    maps = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0, 0, -2., sizes=(n,))[0]
    (3.0 * rp.mask_recon_loss_bis(tgt, maps)).backward()
    dv = _leaves(inp, cuda_device)
    loss = gg.mask_recon_loss_bis(tgt.to(cuda_device), dv["mus"], dv["sigma"], dv["rotx"], dv["roty"], dv["rotz"],
                                  0, 0, -2., size=n)
    (3.0 * loss).backward()
    assert abs(loss.item() - rp.mask_recon_loss_bis(tgt, maps).item()) < 2e-6 * abs(loss.item())
    for k in ("mus", "sigma", "roty"):
        assert_close_norm(dv[k].grad, lv[k].grad, GRAD_RTOL, "loss_bis grad " + k)


@pytest.mark.parametrize("B,K,n", [(4, 16, 256), (2, 6, 64), (2, 40, 48), (1, 3, 100)])
def test_gm_cycle_loss_vs_oracle(gg, cuda_device, B, K, n):
    ia = rp.synth_inputs(B, K, seed=40 + K, angle=10.0)
    ib = {k: v.clone() for k, v in ia.items()}
    g = torch.Generator().manual_seed(2)
    ib["mus"] = ib["mus"] + 0.05 * torch.randn(ib["mus"].shape, generator=g)       # the cycle render is a perturbation
    ib["roty"] = ib["roty"] + 5.0 * torch.randn(ib["roty"].shape, generator=g)
    ib["sigma"] = ib["sigma"] * (1.0 + 0.1 * torch.rand(B, K, 1, 1, generator=g, dtype=torch.float64))
    la, lb = _leaves(ia), _leaves(ib)
    ma = rp.get_landmarks(la["mus"], la["sigma"], la["rotx"], la["roty"], la["rotz"], 0, 0, -2., sizes=(n,))[0]
    mb = rp.get_landmarks(lb["mus"], lb["sigma"], lb["rotx"], lb["roty"], lb["rotz"], 0, 0, -2., sizes=(n,))[0]
    ref = rp.gm_cycle_loss(ma, mb)
    ref.backward()
    da, db = _leaves(ia, cuda_device), _leaves(ib, cuda_device)
    order = ("mus", "sigma", "rotx", "roty", "rotz")
    loss = gg.gm_cycle_loss([da[k] for k in order], [db[k] for k in order], 0, 0, -2., size=n)
    loss.backward()
    assert abs(loss.item() - ref.item()) <= 3e-6 * abs(ref.item())
    for k in ("mus", "sigma", "roty"):
        assert_close_norm(da[k].grad, la[k].grad, GRAD_RTOL, "cycle grad a." + k)
        assert_close_norm(db[k].grad, lb[k].grad, GRAD_RTOL, "cycle grad b." + k)


def test_colorize_both_sources(gg, cuda_device):
    B, K, n = 3, 6, 64
    inp = rp.synth_inputs(B, K, seed=3)
    colors = torch.rand(K, 3, generator=torch.Generator().manual_seed(0))
    maps = rp.get_landmarks(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0, 0, -2., sizes=(n,))[0]
    ref = rp.colorize_landmark_maps(maps, colors)
    d = _dev(inp, cuda_device)
    a = gg.colorize_landmark_maps(maps.to(cuda_device), colors.to(cuda_device))
    assert torch.equal(a.cpu(), ref)                                   # same float32 products and max
    b = gg.colorize_gaussians(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., colors.to(cuda_device), size=n)
    assert_close_norm(b, ref, FWD_RTOL, "colorize from Gaussians")
    u = gg.colorize_landmark_maps(maps.to(cuda_device), colors.to(cuda_device), out="u8_inv")
    expect = 255 - (ref.numpy() * 255).astype(np.uint8)               # mask/infer_masks.py:463
    assert u.dtype == torch.uint8 and np.array_equal(u.cpu().numpy(), expect)


def test_randomised_colorize_shapes(gg, cuda_device):
    """Seeded sweep of the viz composite (mask/mask_gen_dynamic.py:221-231) over K and n (odd sizes, K not a multiple
    of 4): from maps it is exact, from the Gaussians within the forward tolerance, and uint8 frames agree with the
    float ones to the truncation step."""
    rng = torch.Generator().manual_seed(21)
    for case in range(int(os.environ.get("GG_SWEEP_CASES", "100"))):
        B = int(torch.randint(1, 4, (1,), generator=rng))
        K = int(torch.randint(1, 40, (1,), generator=rng))
        n = int(torch.randint(1, 140, (1,), generator=rng))
        inp = rp.synth_inputs(B, K, seed=3000 + case)
        colors = torch.rand(K, 3, generator=rng)
        maps = rp.get_landmarks(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0, 0, -2., sizes=(n,))[0]
        ref = rp.colorize_landmark_maps(maps, colors)
        d = _dev(inp, cuda_device)
        what = f"case {case} (B={B}, K={K}, n={n})"
        a = gg.colorize_landmark_maps(maps.to(cuda_device), colors.to(cuda_device))
        assert torch.equal(a.cpu(), ref), what
        b = gg.colorize_gaussians(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., colors.to(cuda_device), size=n)
        assert_close_norm(b, ref, FWD_RTOL, what)
        u = gg.colorize_gaussians(d["mus"], d["sigma"], d["rotx"], d["roty"], d["rotz"], 0, 0, -2., colors.to(cuda_device),
                                  size=n, out="u8")
        assert u.dtype == torch.uint8 and int((u.cpu().int() - (b.cpu() * 255).int()).abs().max()) <= 1, what


def test_turntable_matches_per_angle_reference_loop(gg, cuda_device):
    K, n = 6, 256
    g = rp.synth_inputs(1, K, seed=56)
    colors = torch.rand(K, 3, generator=torch.Generator().manual_seed(1))
    angles = np.linspace(0.0, 120.0, 7)                                # the reference uses 60 (infer_masks.py:426-428)
    theta0 = torch.tensor([[17.5]])
    out = gg.turntable(g["mus"].to(cuda_device), g["sigma"].to(cuda_device), theta0, angles, colors.to(cuda_device),
                       want_maps=True)
    assert tuple(out["frames"].shape) == (7, n, n, 3) and out["frames"].dtype == torch.uint8
    for f, x in enumerate(angles):                                      # the reference's loop body, one angle at a time
        xr = torch.tensor([[x]], dtype=torch.float32)
        z = torch.zeros_like(xr)
        mu2d, s2d, gm = rp.get_landmarks_infer(g["mus"], g["sigma"], z, theta0[0, 0] + xr * 1., z, 0., 0., -2.)
        np.testing.assert_array_equal(out["mu2d"][f:f + 1].cpu().numpy(), mu2d.numpy())
        np.testing.assert_array_equal(out["sigma2d"][f:f + 1].cpu().numpy(), s2d.numpy())
        assert_close_norm(out["maps"][f:f + 1], gm, FWD_RTOL, "turntable maps")
        frame = 255 - (rp.colorize_landmark_maps(gm, colors).numpy() * 255).astype(np.uint8)
        diff = np.abs(out["frames"][f:f + 1].cpu().numpy().astype(int) - frame.astype(int))
        assert diff.max() <= 1 and (diff > 0).mean() < 1e-3            # uint8 truncation of values 1e-7 apart


def test_interchange_roundtrip(gg, cuda_device, tmp_path):
    K = 12
    g = rp.synth_inputs(1, K, seed=9)
    m_path, s_path = gg.save_reference_gaussians(g["mus"], g["sigma"], str(tmp_path), 3)
    assert m_path.endswith("m_3.npy") and s_path.endswith("sig3d_3.npy")
    assert np.load(m_path).shape == (1, K, 1, 3) and np.load(m_path).dtype == np.float32      # infer_masks.py:447
    assert np.load(s_path).shape == (1, K, 3, 3) and np.load(s_path).dtype == np.float64      # infer_masks.py:448
    mu, sig = gg.load_reference_gaussians(m_path, s_path, cuda_device)
    assert torch.equal(mu.cpu(), g["mus"]) and torch.equal(sig.cpu(), g["sigma"])
    z = torch.zeros(1, 1, device=cuda_device)
    mu2d, s2d, _ = gg.get_landmarks_infer(mu, sig, z, z + 30., z, 0., 0., -2.)
    feed = gg.decoder_feed(mu2d, s2d)
    assert feed["gen/genlandmarks/mu2d:0"].shape == (1, K, 2) and feed["gen/genlandmarks/sigma2d:0"].dtype == np.float32


# --------------------------------------------------- Gaussian construction / deformation
def test_sigma_from_frame_vs_oracle(gg, cuda_device):
    g = torch.Generator().manual_seed(11)
    N, K = 3, 16
    raw = [torch.tanh(torch.randn(N, K, 3, generator=g)), torch.tanh(torch.randn(N, K, 3, generator=g)),
           torch.randn(N, K, 3, generator=g)]
    ref_in = [t.clone().requires_grad_(True) for t in raw]
    ref = rp.sigma_from_frame(*ref_in)
    w = torch.randn(ref.shape, generator=g, dtype=torch.float64)
    (ref * w).sum().backward()
    dev_in = [t.to(cuda_device).requires_grad_(True) for t in raw]
    out = gg.sigma_from_frame(*dev_in)
    assert out.dtype == torch.float64 and out.shape == ref.shape
    assert_close_norm(out, ref, 1e-6, "sigma_from_frame (float32 arithmetic, 1-2 ulp of sigmoid/rsqrt)")
    (out * w.to(cuda_device)).sum().backward()
    for a, b, nm in zip(dev_in, ref_in, ("v0", "v1", "u")):
        assert_close_norm(a.grad, b.grad, GRAD_RTOL, "sigma_from_frame grad " + nm)


def test_deform_vs_oracle(gg, cuda_device):
    B, K = 5, 12
    g = torch.Generator().manual_seed(21)
    th = lambda *s: torch.tanh(torch.randn(*s, generator=g))
    cann = rp.synth_inputs(1, K, seed=56)
    heads = [th(B, K * 3), th(B, K * 3), th(B, K), th(B, K), th(B, K), th(B, 1)]
    wm, ws, wt = torch.randn(B, K, 1, 3, generator=g), torch.randn(B, K, 3, 3, generator=g, dtype=torch.float64), torch.randn(B, 1, generator=g)

    def run(fn, dev):
        cm = cann["mus"].detach().clone().to(dev).requires_grad_(True)
        cs = cann["sigma"].detach().clone().to(dev).requires_grad_(True)
        hs = [h.detach().clone().to(dev).requires_grad_(True) for h in heads]
        mus, sig, theta = fn(*hs, cm, cs)
        ((mus * wm.to(dev)).sum() + (sig * ws.to(dev)).sum() + (theta * wt.to(dev)).sum()).backward()
        return (mus, sig, theta), [cm.grad, cs.grad] + [h.grad for h in hs]

    (rm, rs, rt), gref = run(rp.deform, torch.device("cpu"))
    (om, os_, ot), gout = run(gg.deform, cuda_device)
    assert om.shape == rm.shape and om.dtype == torch.float32 and os_.dtype == torch.float64 and ot.shape == rt.shape
    assert torch.equal(om.cpu(), rm.detach()) and torch.equal(ot.cpu(), rt.detach())
    assert_close_norm(os_, rs, 1e-9, "deform sigmas")       # float64 chain; CPU vs GPU SVD differ at round-off
    for a, b, nm in zip(gout, gref, ["cannmu", "cannsigma", "mu_t", "scale", "rx", "ry", "rz", "theta"]):
        assert_close_norm(a, b, GRAD_RTOL, "deform grad " + nm)


def test_config3_full_cuda_pipeline(gg, cuda_device):
    """Config 3 end to end on the CUDA path: frame -> Sigma_c, deformation, render, fused loss; gradients reach the
    canonical frame parameters and the deformation heads.  Checked against the oracle's same chain."""
    B, K, n = 6, 16, 128
    g = torch.Generator().manual_seed(33)
    th = lambda *s: torch.tanh(torch.randn(*s, generator=g))
    frame = [th(1, K, 3), th(1, K, 3), torch.randn(1, K, 3, generator=g)]
    cmu = th(1, K, 1, 3)
    heads = [th(B, K * 3), th(B, K * 3), th(B, K), th(B, K), th(B, K), th(B, 1)]
    tgt = rp.synth_target_masks(B, n, seed=4)

    def run(mod, dev, loss_fn):
        fr = [t.detach().clone().to(dev).requires_grad_(True) for t in frame]
        cm = cmu.detach().clone().to(dev).requires_grad_(True)
        hs = [h.detach().clone().to(dev).requires_grad_(True) for h in heads]
        csig = mod.sigma_from_frame(*fr)
        mus, sig, theta = mod.deform(*hs, cm, csig)
        loss = loss_fn(mus, sig, torch.zeros_like(theta), theta, torch.zeros_like(theta), dev)
        loss.backward()
        return loss.detach(), [t.grad for t in fr] + [cm.grad] + [h.grad for h in hs]

    ref_loss = lambda mus, sig, rx, ry, rz, dev: rp.mask_recon_loss_bis(
        tgt, rp.get_landmarks(mus, sig, rx, ry, rz, 0, 0, -2., sizes=(n,))[0])
    out_loss = lambda mus, sig, rx, ry, rz, dev: gg.mask_recon_loss_bis(tgt.to(dev), mus, sig, rx, ry, rz, 0, 0, -2., size=n)
    lr, gr = run(rp, torch.device("cpu"), ref_loss)
    lo, go = run(gg, cuda_device, out_loss)
    assert abs(lo.item() - lr.item()) <= 1e-5 * abs(lr.item())
    names = ["v0", "v1", "u", "cannmu", "mu_t", "scale", "rx", "ry", "rz", "theta"]
    for a, b, nm in zip(go, gr, names):
        assert_close_norm(a, b, GRAD_RTOL, "C3 pipeline grad " + nm)


# ------------------------------------------------------------ forward-differenced mask kernels / PDL
@pytest.fixture
def restore_modes(gg):
    yield
    gg.set_fast_mask(True)
    gg.set_pdl(True)


def _mask_and_grads(gg, d, n, gm):
    lv = {k: d[k].clone().requires_grad_(True) for k in ("mus", "sigma", "roty")}
    m = gg.render_silhouette(lv["mus"], lv["sigma"], d["rotx"], lv["roty"], d["rotz"], 0, 0, -2., size=n)
    m.backward(gm)
    return m.detach(), {k: v.grad for k, v in lv.items()}


@pytest.mark.parametrize("n", [256, 128, 100, 64, 33, 32])
def test_fast_and_direct_mask_arithmetic_vs_oracle(gg, cuda_device, restore_modes, n):
    """Mask-only launches: the forward-differenced kernels (default) and the one-exponential-per-pixel kernels
    against the float64 oracle, forward and gradients, on the reference's activation ranges."""
    B, K = 6, 16
    inp = rp.synth_inputs(B, K, seed=40 + n, angle=30.0)
    lv = _leaves(inp)
    ref = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0, 0, -2., sizes=(n,))[0].sum(-1)
    gm = torch.randn(B, n, n, generator=torch.Generator().manual_seed(n))
    (ref * gm).sum().backward()
    d = _dev(inp, cuda_device)
    errs = {}
    for fast in (True, False):
        gg.set_fast_mask(fast)
        m, g = _mask_and_grads(gg, d, n, gm.to(cuda_device))
        errs[fast] = max_rel(m, ref)
        assert_close_norm(m, ref, FWD_RTOL if fast else 2e-6, f"mask fast={fast} n={n}")
        for k in ("mus", "sigma", "roty"):
            assert_close_norm(g[k], lv[k].grad, GRAD_RTOL, f"grad {k} fast={fast} n={n}")
    assert errs[True] < 3e-6, errs        # measured ~6e-7: an order of magnitude inside the 1e-5 tolerance


def test_steep_gaussians_fall_back_to_direct_evaluation(gg, cuda_device, restore_modes):
    """Sub-pixel-thin, very elongated and far off-screen Gaussians: the recurrence's steepness guard must route
    them to the direct form -- finite everywhere and equal to the oracle, forward and backward."""
    B, K, n = 2, 6, 64
    mus = torch.tensor([[[0.0, 0, 0], [0.4, -0.3, 0.1], [2.5, 0, 0], [0, -2.5, 0.2], [-0.2, 0.2, 0], [0.1, 0.1, 0.1]]] * B)
    mus = mus.view(B, K, 1, 3).float()
    sig = torch.eye(3, dtype=torch.float64).view(1, 1, 3, 3).repeat(B, K, 1, 1) * 0.02
    sig[:, 0] *= 1e-3                                    # far thinner than a pixel
    sig[:, 1, 0, 0] *= 1e-3                              # needle along y
    sig[:, 4] *= 4.0
    sig[:, 5, 0, 1] = sig[:, 5, 1, 0] = 0.015            # tilted
    rot = (torch.tensor([[20.0], [-35.0]]), torch.tensor([[10.0], [160.0]]), torch.tensor([[0.0], [5.0]]))
    lv = [t.clone().requires_grad_(True) for t in (mus, sig)]
    ref = rp.get_landmarks(lv[0], lv[1], rot[0], rot[1], rot[2], 0, 0, -2., sizes=(n,))[0].sum(-1)
    gm = torch.randn(B, n, n, generator=torch.Generator().manual_seed(2))
    (ref * gm).sum().backward()
    for fast in (True, False):
        gg.set_fast_mask(fast)
        dv = [t.to(cuda_device).requires_grad_(True) for t in (mus, sig)]
        r = [t.to(cuda_device) for t in rot]
        m = gg.render_silhouette(dv[0], dv[1], r[0], r[1], r[2], 0, 0, -2., size=n)
        assert torch.isfinite(m).all()
        assert_close_norm(m, ref, FWD_RTOL, f"steep mask fast={fast}")
        m.backward(gm.to(cuda_device))
        assert torch.isfinite(dv[0].grad).all() and torch.isfinite(dv[1].grad).all()
        assert_close_norm(dv[0].grad, lv[0].grad, GRAD_RTOL, f"steep dmus fast={fast}")
        assert_close_norm(dv[1].grad, lv[1].grad, GRAD_RTOL, f"steep dsigma fast={fast}")


def test_programmatic_dependent_launch_changes_nothing(gg, cuda_device, restore_modes):
    """Every kernel waits on its predecessor before touching memory: 40 back-to-back steps on changing inputs give
    bit-identical results with PDL on and off (a missing wait would show up as a race here)."""
    from gaussigan_b200.plan import SilhouetteLossStep, SilhouetteStep
    B, K, n = 16, 16, 128
    sets = []
    for i in range(4):
        d = _dev(rp.synth_inputs(B, K, seed=900 + i), cuda_device)
        sets.append((d["mus"].reshape(B, K, 3).contiguous(), d["sigma"],
                     *(d[k].reshape(B).contiguous() for k in ("rotx", "roty", "rotz")),
                     torch.randn(B, n, n, device=cuda_device),
                     (torch.rand(B, n, n, device=cuda_device) * 255).to(torch.uint8)))
    out = {}
    for pdl in (True, False):
        gg.set_pdl(pdl)
        step, fstep = SilhouetteStep(B, K, n, cuda_device), SilhouetteLossStep(B, K, n, cuda_device)
        acc = []
        for it in range(40):
            mus, sig, rx, ry, rz, gm, tgt = sets[it % 4]
            step.forward(mus, sig, rx, ry, rz)
            g = [t.clone() for t in step.backward(gm)]
            fstep.bind(mus, sig, rx, ry, rz)
            fg = [t.clone() for t in fstep.loss_backward(tgt)]
            acc.append([step.mask.clone()] + g + fg + [fstep.loss_partials.clone()])
        torch.cuda.synchronize()
        out[pdl] = acc
    for a, b in zip(out[True], out[False]):
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_batches_beyond_one_grid_dimension(gg, cuda_device):
    """More than 65535 batch elements in one call (beyond one grid dimension): tiny images, the
    first and last batch elements against the oracle, forward, fused loss and gradients."""
    from gaussigan_b200.plan import SilhouetteLossStep
    B, K, n = 65535 + 9, 2, 8
    base = rp.synth_inputs(16, K, seed=5)
    rep = (B + 15) // 16
    inp = {k: v.repeat(rep, *([1] * (v.dim() - 1)))[:B].contiguous() for k, v in base.items()}
    inp["roty"] = (inp["roty"] + torch.arange(B, dtype=torch.float32).view(B, 1) * 1e-3).contiguous()   # all different
    d = _dev(inp, cuda_device)
    lv = {k: d[k].clone().requires_grad_(True) for k in ("mus", "sigma", "roty")}
    m = gg.render_silhouette(lv["mus"], lv["sigma"], d["rotx"], lv["roty"], d["rotz"], 0, 0, -2., size=n)
    gm = torch.randn(B, n, n, generator=torch.Generator().manual_seed(1))
    m.backward(gm.to(cuda_device))
    sel = torch.cat([torch.arange(0, 5), torch.arange(65530, 65540), torch.arange(B - 5, B)])
    sub_in = {k: v[sel].clone().requires_grad_(True) for k, v in inp.items()}
    ref = rp.get_landmarks(sub_in["mus"], sub_in["sigma"], sub_in["rotx"], sub_in["roty"], sub_in["rotz"], 0, 0, -2.,
                           sizes=(n,))[0].sum(-1)
    (ref * gm[sel]).sum().backward()
    assert_close_norm(m[sel.to(cuda_device)], ref, FWD_RTOL, "big-batch mask")
    for k in ("mus", "sigma", "roty"):
        assert_close_norm(lv[k].grad[sel.to(cuda_device)], sub_in[k].grad, GRAD_RTOL, f"big-batch grad {k}")
    # fused loss kernel on the same batch: loss equals the mean over all images of the unfused residual
    tgt = (torch.rand(B, n, n, generator=torch.Generator().manual_seed(2)) * 255).to(torch.uint8)
    step = SilhouetteLossStep(B, K, n, cuda_device)
    step.bind(d["mus"].reshape(B, K, 3).contiguous(), d["sigma"], *(d[k].reshape(B).contiguous() for k in ("rotx", "roty", "rotz")))
    step.loss_backward(tgt.to(cuda_device))
    t01 = ((tgt.float() / 127.5 - 1.0) + 1.0) * 0.5
    want = float((t01.to(cuda_device) - m.detach()).abs().double().mean())
    assert abs(step.loss() - want) <= 1e-6 * abs(want)


@pytest.mark.parametrize("pinned,B,n", [(True, 3, 64), (False, 3, 64), (True, 1, 33)])
def test_host_buffer_step_many_gaussians_and_pageable_buffers(gg, cuda_device, pinned, B, n):
    """K > 64 takes the five-launch path (splat, loss kernel, splat backward); pageable host buffers take the
    copy path instead of the in-place (zero-copy) access of pinned ones.  Same numbers as the oracle either way."""
    from gaussigan_b200.host import HostInputs, HostSilhouetteStep
    K = 80                                              # B = 1, n = 33: a pixel count that is not a multiple of 4
    inp = rp.synth_inputs(B, K, seed=123)
    tgt = rp.synth_target_masks(B, n, seed=8)
    mask_ref, loss_ref, dmu_ref, dsig_ref, drot_ref = rp.render_loss_and_grads(inp, tgt, n)
    host = HostSilhouetteStep(B, K, n, cuda_device, depth=2, want_mask=True)
    assert host.LAUNCHES_PER_STEP == 5
    h = host.make_host_inputs(inp, tgt)
    if not pinned:
        h = HostInputs(*(t.clone() for t in h))          # plain pageable copies
        assert not h.mus.is_pinned()
    res = host.submit(h).wait()
    torch.cuda.synchronize()
    assert abs(res.loss - float(loss_ref)) <= 2e-6 * abs(float(loss_ref))
    assert_close_norm(res.mask, mask_ref, FWD_RTOL, "host K=80 mask")
    assert_close_norm(res.dmus, dmu_ref.reshape(B, K, 3), GRAD_RTOL, "host K=80 dmus")
    assert_close_norm(res.dsigma, dsig_ref, GRAD_RTOL, "host K=80 dsigma")
    assert_close_norm(res.drot[:, 1], drot_ref.reshape(B), GRAD_RTOL, "host K=80 droty")


def test_svd3_kernel_vs_torch_svd(gg, cuda_device):
    """The one-sided Jacobi 3x3 SVD kernel (mask/mask_gen_dynamic.py:430 without a host sync): singular values,
    reconstruction and orthogonality to float64 round-off on the reference's slightly asymmetric covariances and on
    general matrices; gradients of a sign-invariant function of (U, S) equal those through torch.linalg.svd; and
    deform() gives the same Sigma_b and gradients with either factorisation."""
    g = torch.Generator().manual_seed(3)
    K = 16
    th = lambda *s: torch.tanh(torch.randn(*s, generator=g))
    sig_c = rp.sigma_from_frame(th(1, K, 3), th(1, K, 3), torch.randn(1, K, 3, generator=g)).reshape(K, 3, 3)
    general = torch.randn(8, 3, 3, generator=g, dtype=torch.float64)
    for mats in (sig_c, general):
        A = mats.to(cuda_device)
        U, S = gg.svd3(A)
        Ut, St, Vh = torch.linalg.svd(mats)
        assert (S.cpu() - St).abs().max() < 1e-13 and bool((S[:, :-1] >= S[:, 1:]).all())
        assert (U.transpose(-1, -2) @ U - torch.eye(3, dtype=torch.float64, device=cuda_device)).abs().max() < 1e-13
        # same left singular vectors up to sign
        dots = (U.cpu() * Ut).sum(-2).abs()
        assert (dots - 1).abs().max() < 1e-10
    # gradient of a sign-invariant scalar: sum W * (U diag(sqrt(S)) M diag(sqrt(S)) U^T)
    W = torch.randn(K, 3, 3, generator=g, dtype=torch.float64)
    M = torch.diag(torch.tensor([1.3, 0.7, 2.1], dtype=torch.float64))

    def f(U_, S_, W_, M_):
        D = torch.diag_embed(S_.sqrt())
        return (W_ * (U_ @ D @ M_ @ D @ U_.transpose(-1, -2))).sum()

    a1 = sig_c.to(cuda_device).clone().requires_grad_(True)
    f(*gg.svd3(a1), W.to(cuda_device), M.to(cuda_device)).backward()
    a2 = sig_c.clone().requires_grad_(True)
    U2, S2, _ = torch.linalg.svd(a2)
    f(U2, S2, W, M).backward()
    assert_close_norm(a1.grad, a2.grad, 1e-9, "svd3 gradient")
    # deform() with either factorisation
    B = 5
    heads = [th(B, K * 3), th(B, K * 3), th(B, K), th(B, K), th(B, K), th(B, 1)]
    cmu = th(1, K, 1, 3)
    outs = {}
    for mode in ("cuda", "torch"):
        cs = sig_c.reshape(1, K, 3, 3).to(cuda_device).clone().requires_grad_(True)
        hs = [h.to(cuda_device).clone().requires_grad_(True) for h in heads]
        if mode == "cuda":
            mus, sg, theta = gg.deform(*hs, cmu.to(cuda_device), cs)
        else:       # the test's own cuSOLVER factorisation handed to the same kernels (the product has no such route)
            Ut, St, _ = torch.linalg.svd(cs.reshape(K, 3, 3), full_matrices=True)
            mus, sg, theta = gg.deform(*hs, cmu.to(cuda_device), cs, factors=(Ut, St))
        (sg * torch.randn(sg.shape, generator=torch.Generator().manual_seed(1), dtype=torch.float64).to(cuda_device)).sum().backward()
        outs[mode] = (sg.detach(), cs.grad, [h.grad for h in hs[:5]])
    assert_close_norm(outs["cuda"][0], outs["torch"][0], 1e-11, "Sigma_b cuda vs torch svd")
    assert_close_norm(outs["cuda"][1], outs["torch"][1], 1e-8, "d cannsigma")
    for a, b in zip(outs["cuda"][2], outs["torch"][2]):
        assert_close_norm(a, b, 1e-6, "head gradients")


def test_config3_step_is_graph_capturable(gg, cuda_device):
    """The whole dynamic-object step (frame -> Sigma_c, SVD, deformation, render, fused loss, backward) never touches
    the host, so it can be captured in one CUDA graph; replays give the eager gradients bit for bit."""
    B, K, n = 4, 16, 64
    g = torch.Generator().manual_seed(12)
    th = lambda *s: torch.tanh(torch.randn(*s, generator=g))
    leaves = [t.to(cuda_device).requires_grad_(True) for t in
              (th(1, K, 3), th(1, K, 3), torch.randn(1, K, 3, generator=g), th(1, K, 1, 3),
               th(B, K * 3), th(B, K * 3), th(B, K), th(B, K), th(B, K), th(B, 1))]
    tgt = rp.synth_target_masks(B, n, seed=9).to(cuda_device)

    def step():
        for t in leaves:
            t.grad.zero_()
        csig = gg.sigma_from_frame(*leaves[:3])
        mus, sig, theta = gg.deform(*leaves[4:], leaves[3], csig)
        z = torch.zeros_like(theta)
        loss = gg.mask_recon_loss_bis(tgt, mus, sig, z, theta, z, 0, 0, -2., size=n)
        loss.backward()
        return loss

    for t in leaves:
        t.grad = torch.zeros_like(t)
    side = torch.cuda.Stream(cuda_device)
    side.wait_stream(torch.cuda.current_stream(cuda_device))
    with torch.cuda.stream(side):
        for _ in range(3):
            eager_loss = step().detach().clone()
    torch.cuda.current_stream(cuda_device).wait_stream(side)
    torch.cuda.synchronize()
    eager = [t.grad.clone() for t in leaves]
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss = step()
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(loss.detach(), eager_loss)
    for a, b in zip(leaves, eager):
        assert torch.equal(a.grad, b)


def test_randomised_shapes_and_scales_fast_mask_path(gg, cuda_device, restore_modes):
    """Seeded sweep over image sizes, Gaussian sizes from about a pixel to larger than the frame (axis variances
    3e-4 .. 0.6; the reference's own range is 0.01 .. 0.51, far sub-pixel Gaussians are beyond what a float32 raster
    can resolve to 1e-5 in either arithmetic), off-screen centres, all three Euler angles, non-default translation
    and focal: the forward-differenced kernels (with their steepness guard) stay inside the path's tolerances."""
    rng = torch.Generator().manual_seed(2024)
    worst_f, worst_g = 0.0, 0.0
    for case in range(int(os.environ.get("GG_SWEEP_CASES", "120"))):      # larger values for a soak run
        B, K = 2, int(torch.randint(1, 9, (1,), generator=rng))
        n = int(torch.randint(9, 200, (1,), generator=rng))
        mus = ((torch.rand(B, K, 1, 3, generator=rng) - 0.5) * torch.tensor([3.0, 3.0, 1.0])).float()
        # covariances: random rotation, axis variances log-uniform in [3e-4, 0.6]
        q, _ = torch.linalg.qr(torch.randn(B, K, 3, 3, generator=rng, dtype=torch.float64))
        lo, hi = torch.log(torch.tensor(3e-4, dtype=torch.float64)), torch.log(torch.tensor(0.6, dtype=torch.float64))
        var = torch.exp(torch.rand(B, K, 3, generator=rng, dtype=torch.float64) * (hi - lo) + lo)
        sig = (q * var[..., None, :]) @ q.transpose(-1, -2)
        rot = [(torch.rand(B, 1, generator=rng) - 0.5) * a for a in (40.0, 360.0, 40.0)]
        t = (float(torch.rand(1, generator=rng) - 0.5) * 0.4, float(torch.rand(1, generator=rng) - 0.5) * 0.4,
             -2.0 - float(torch.rand(1, generator=rng)))
        focal = 0.8 + float(torch.rand(1, generator=rng)) * 0.6
        lv = [x.clone().requires_grad_(True) for x in (mus, sig, rot[1])]
        ref = rp.get_landmarks(lv[0], lv[1], rot[0], lv[2], rot[2], *t, focal, sizes=(n,))[0].sum(-1)
        if not torch.isfinite(ref).all() or float(ref.detach().abs().max()) < 1e-6:
            continue                                         # degenerate geometry / empty frame: nothing to compare
        gm = torch.randn(B, n, n, generator=rng)
        (ref * gm).sum().backward()
        dv = [x.to(cuda_device).requires_grad_(True) for x in (mus, sig, rot[1])]
        m = gg.render_silhouette(dv[0], dv[1], rot[0].to(cuda_device), dv[2], rot[2].to(cuda_device), *t, focal, size=n)
        m.backward(gm.to(cuda_device))
        assert torch.isfinite(m).all(), f"case {case}"
        ef = max_rel(m, ref)
        worst_f = max(worst_f, ef)
        assert ef <= FWD_RTOL, f"case {case} (n={n}, K={K}): forward {ef:.2e}"
        for a, b, nm in zip(dv, lv, ("mus", "sigma", "roty")):
            if float(b.grad.abs().max()) < 1e-12:
                continue
            eg = max_rel(a.grad, b.grad)
            worst_g = max(worst_g, eg)
            assert eg <= GRAD_RTOL, f"case {case} (n={n}, K={K}): grad {nm} {eg:.2e}"
    print(f"worst forward {worst_f:.2e}, worst gradient {worst_g:.2e}")


@pytest.mark.parametrize("K,sizes,B", [
    (16, (32, 64, 128, 256), 3), (4, (8, 24), 2), (8, (40, 64), 2), (32, (48,), 2), (64, (72, 16), 1), (128, (32,), 1),
    (6, (32, 64, 128, 256), 2), (12, (32, 64, 128), 2), (5, (40, 64), 2), (7, (24,), 3), (1, (16,), 2), (2, (64,), 2),
    (20, (56,), 2), (63, (16,), 1),
])
def test_nhwc_maps_backward_forward_differenced_vs_oracle_and_direct(gg, cuda_device, restore_modes, K, sizes, B):
    """NHWC maps + silhouette gradients through the forward-differenced bulk-copy kernel (gg_splat_tma.cu: quads of
    Gaussians for K % 4 == 0, pairs for every other K including odd ones and the reference's own 6 and 12; any n that
    is a multiple of 8, strips of one warp slice in different rows at small n, partial last slices) against the
    oracle's autograd, and against the direct kernels (gg.set_fast_mask(False)).  The forward of the non-power-of-two
    K goes through the flat 128-bit-store kernel."""
    inp = rp.synth_inputs(B, K, seed=700 + K, angle=45.0)
    # one needle narrower than a pixel at the finest scale (the kernel's direct fall-back for its unit) and one blob
    # far outside the frame (above it: the yaw leaves y alone, so it stays in front of the camera)
    inp["sigma"][0, 0] = torch.diag(torch.tensor([3e-4, 0.2, 0.05], dtype=torch.float64))
    inp["mus"][B - 1, K - 1, 0] = torch.tensor([0.1, 2.5, 0.1])
    lv = _leaves(inp)
    ref = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0, 0, -2., sizes=sizes)
    w = weights_like([r.shape for r in ref], 31)
    wm = weights_like([(B, sizes[-1], sizes[-1], 1)], 32)[0]
    (sum((r * q).sum() for r, q in zip(ref, w)) + (ref[-1].sum(-1, keepdim=True) * wm).sum()).backward()

    def run(fast):
        gg.set_fast_mask(fast)
        dv = _leaves(inp, cuda_device)
        out, mask = gg.get_landmarks_with_mask(dv["mus"], dv["sigma"], dv["rotx"], dv["roty"], dv["rotz"], 0, 0, -2.,
                                               sizes=sizes)
        for o, r in zip(out, ref):
            assert_close_norm(o, r, FWD_RTOL, f"K={K} maps n={r.shape[1]}")
        assert_close_norm(mask, ref[-1].detach().sum(-1, keepdim=True), FWD_RTOL, f"K={K} silhouette")
        (sum((o * q.to(cuda_device)).sum() for o, q in zip(out, w)) + (mask * wm.to(cuda_device)).sum()).backward()
        return {k: dv[k].grad for k in ("mus", "sigma", "rotx", "roty", "rotz")}

    fast, direct, again = run(True), run(False), run(True)
    for k in fast:
        assert torch.equal(fast[k], again[k]), f"K={K} grad {k}: not bit-reproducible"      # fixed reduction order
        assert_close_norm(fast[k], lv[k].grad, GRAD_RTOL, f"K={K} grad {k} (forward-differenced)")
        assert_close_norm(direct[k], lv[k].grad, GRAD_RTOL, f"K={K} grad {k} (direct)")
        assert_close_norm(fast[k], direct[k], GRAD_RTOL, f"K={K} grad {k} fast vs direct")


def test_randomised_nhwc_maps_forward_and_backward(gg, cuda_device):
    """Seeded sweep of the NHWC maps kernels (quad / pair / flat forwards, quad- and pair-unit forward-differenced
    backward): K 1..24, n a multiple of 8 up to 160, Gaussians from about a pixel to larger than the frame, off-screen
    centres, all three Euler angles, with and without a silhouette gradient on top of the maps gradient."""
    rng = torch.Generator().manual_seed(77)
    worst_f, worst_g = 0.0, 0.0
    for case in range(int(os.environ.get("GG_SWEEP_CASES", "100"))):      # larger values for a soak run
        B, K = 2, int(torch.randint(1, 25, (1,), generator=rng))
        n = 8 * int(torch.randint(1, 21, (1,), generator=rng))
        mus = ((torch.rand(B, K, 1, 3, generator=rng) - 0.5) * torch.tensor([3.0, 3.0, 1.0])).float()
        q, _ = torch.linalg.qr(torch.randn(B, K, 3, 3, generator=rng, dtype=torch.float64))
        lo, hi = torch.log(torch.tensor(3e-4, dtype=torch.float64)), torch.log(torch.tensor(0.6, dtype=torch.float64))
        var = torch.exp(torch.rand(B, K, 3, generator=rng, dtype=torch.float64) * (hi - lo) + lo)
        sig = (q * var[..., None, :]) @ q.transpose(-1, -2)
        rot = [(torch.rand(B, 1, generator=rng) - 0.5) * a for a in (40.0, 360.0, 40.0)]
        with_mask = bool(torch.randint(0, 2, (1,), generator=rng))
        lv = [x.clone().requires_grad_(True) for x in (mus, sig, rot[1])]
        ref = rp.get_landmarks(lv[0], lv[1], rot[0], lv[2], rot[2], 0, 0, -2.5, sizes=(n,))[0]
        if not torch.isfinite(ref).all() or float(ref.detach().abs().max()) < 1e-6 or float(ref.detach().max()) > 1.5:
            continue                       # degenerate geometry (a blob through the camera plane) / empty frame
        w = torch.randn(B, n, n, K, generator=rng)
        wm = torch.randn(B, n, n, 1, generator=rng)
        ((ref * w).sum() + ((ref.sum(-1, keepdim=True) * wm).sum() if with_mask else 0.0)).backward()
        dv = [x.to(cuda_device).requires_grad_(True) for x in (mus, sig, rot[1])]
        out, mask = gg.get_landmarks_with_mask(dv[0], dv[1], rot[0].to(cuda_device), dv[2], rot[2].to(cuda_device),
                                               0, 0, -2.5, sizes=(n,))
        ef = max(max_rel(out[0], ref), max_rel(mask, ref.detach().sum(-1, keepdim=True)))
        worst_f = max(worst_f, ef)
        assert ef <= FWD_RTOL, f"case {case} (n={n}, K={K}): forward {ef:.2e}"
        ((out[0] * w.to(cuda_device)).sum() + ((mask * wm.to(cuda_device)).sum() if with_mask else 0.0)).backward()
        for a, b, nm in zip(dv, lv, ("mus", "sigma", "roty")):
            if float(b.grad.abs().max()) < 1e-12:
                continue
            eg = max_rel(a.grad, b.grad)
            worst_g = max(worst_g, eg)
            assert eg <= GRAD_RTOL, f"case {case} (n={n}, K={K}, mask={with_mask}): grad {nm} {eg:.2e}"
    print(f"worst forward {worst_f:.2e}, worst gradient {worst_g:.2e}")


def test_second_device_in_the_same_process(gg, cuda_device):
    """Tensors on cuda:1 while cuda:0 is PyTorch's current device (and the other way round): the library follows the
    tensors' device on every call.  Needs two GPUs."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    B, K, n = 3, 8, 64
    inp = rp.synth_inputs(B, K, seed=3)
    outs = []
    for dev_index, current in ((0, 0), (1, 0), (0, 1), (1, 1)):
        torch.cuda.set_device(current)
        d = {k: v.to(f"cuda:{dev_index}") for k, v in inp.items()}
        lv = {k: d[k].clone().requires_grad_(True) for k in ("mus", "sigma", "roty")}
        m = gg.render_silhouette(lv["mus"], lv["sigma"], d["rotx"], lv["roty"], d["rotz"], 0, 0, -2., size=n)
        m.sum().backward()
        torch.cuda.synchronize(dev_index)
        assert m.device.index == dev_index
        outs.append((m.cpu(), lv["mus"].grad.cpu(), lv["sigma"].grad.cpu()))
    torch.cuda.set_device(0)
    for o in outs[1:]:
        for a, b in zip(o, outs[0]):
            assert torch.equal(a, b)


@pytest.mark.parametrize("B,K,n", [(5, 16, 96), (1, 1, 8), (3, 6, 33), (2, 12, 40), (4, 37, 64), (2, 64, 17)])
def test_planned_dynamic_object_step_equals_autograd_chain(gg, cuda_device, B, K, n):
    """DynamicObjectStep (config 3 as a fixed launch sequence, graph-captured) against the autograd chain
    sigma_from_frame -> deform -> mask_recon_loss_bis, which test_config3_full_cuda_pipeline pins to the oracle."""
    from gaussigan_b200.plan import DynamicObjectStep
    g = torch.Generator().manual_seed(77 + K)
    th = lambda *s: torch.tanh(torch.randn(*s, generator=g)).to(cuda_device)
    v0, v1, u_raw, cmu = th(K, 3), th(K, 3), torch.randn(K, 3, generator=g).to(cuda_device), th(K, 3)
    mu_t, scale, ang, theta_raw = th(B, K, 3), th(B, K, 3), th(B, K, 3), th(B)
    tgt = (torch.rand(B, n, n, generator=g) * 255).to(torch.uint8).to(cuda_device)
    step = DynamicObjectStep(B, K, n, cuda_device)
    step.bind(v0, v1, u_raw, cmu, mu_t, scale, ang, theta_raw)
    step.capture(tgt)
    for t in step.grads.values():
        t.zero_()
    step.graph.replay()
    torch.cuda.synchronize()
    # autograd chain on the same numbers
    lv = [t.clone().requires_grad_(True) for t in (v0, v1, u_raw, cmu, mu_t, scale, ang, theta_raw)]
    csig = gg.sigma_from_frame(lv[0].view(1, K, 3), lv[1].view(1, K, 3), lv[2].view(1, K, 3))
    mus, sig, theta = gg.deform(lv[4].reshape(B, K * 3), lv[5].reshape(B, K * 3), lv[6][..., 0], lv[6][..., 1], lv[6][..., 2],
                                lv[7].view(B, 1), lv[3].view(1, K, 1, 3), csig)
    z = torch.zeros_like(theta)
    loss = gg.mask_recon_loss_bis(tgt.float() / 127.5 - 1.0, mus, sig, z, theta, z, 0, 0, -2., size=n)
    loss.backward()
    assert abs(step.loss() - float(loss)) <= 1e-6 * abs(float(loss))
    for name, ref in zip(("v0", "v1", "u_raw", "cannmu", "mu_t", "scale", "ang", "theta_raw"), lv):
        assert_close_norm(step.grads[name], ref.grad, 1e-5, "planned dynamic step grad " + name)


def test_p2p_shared_grad_reducer_single_rank(gg, cuda_device):
    """gg_shared_grad_post_p2p / gg_shared_grad_wait_p2p with a one-rank group (the 2-rank run incl. a delayed and a
    timed-out peer is tests/test_dist_gpu.py; 8 ranks: tools/p2p_allreduce_check.py, results in profiles/): local batch
    sum in float64 in a fixed order, slot/flag protocol over many calls (both slot sets, epochs), CUDA-graph replay."""
    import torch.distributed as dist
    from gaussigan_b200.dist import P2PSharedGradReducer
    created = False
    if not dist.is_initialized():
        os_env = __import__("os").environ
        os_env.setdefault("MASTER_ADDR", "127.0.0.1")
        os_env.setdefault("MASTER_PORT", "29611")
        try:
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=cuda_device)
        except Exception as exc:                                   # pragma: no cover
            pytest.skip(f"cannot create a one-rank NCCL group here: {exc!r}")
        created = True
    try:
        try:
            K, B = 16, 37
            red = P2PSharedGradReducer(K, cuda_device)
        except Exception as exc:                                   # pragma: no cover
            pytest.skip(f"symmetric memory unavailable: {exc!r}")
        g = torch.Generator().manual_seed(4)
        for it in range(7):
            dmus = torch.randn(B, K, 3, generator=g).to(cuda_device)
            dsig = torch.randn(B, K, 3, 3, generator=g, dtype=torch.float64).to(cuda_device)
            a, b = red(dmus, dsig)
            red.check()
            ra = torch.zeros(K, 3, dtype=torch.float64, device=cuda_device)
            rb = torch.zeros(K, 3, 3, dtype=torch.float64, device=cuda_device)
            per = (B + 3) // 4                                      # the kernel's order: 4 contiguous batch slices,
            for s0 in range(0, B, per):                             # each ascending, then the slices in order
                pa = torch.zeros_like(ra)
                pb = torch.zeros_like(rb)
                for i in range(s0, min(B, s0 + per)):
                    pa += dmus[i].double()
                    pb += dsig[i]
                ra += pa
                rb += pb
            assert torch.equal(a, ra) and torch.equal(b, rb)
        # the two halves separately, with other work in between (post ... wait)
        red.post(dmus, dsig)
        filler = torch.randn(1 << 20, device=cuda_device).sin().sum()
        a2, b2 = red.wait()
        red.check()
        assert torch.equal(a2, ra) and torch.equal(b2, rb) and bool(torch.isfinite(filler))
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(cuda_device)
        side.wait_stream(torch.cuda.current_stream(cuda_device))
        with torch.cuda.stream(side):
            red(dmus, dsig)
        torch.cuda.current_stream(cuda_device).wait_stream(side)
        torch.cuda.synchronize()
        with torch.cuda.graph(graph):
            red(dmus, dsig)
        for _ in range(5):
            graph.replay()
        red.check()
        assert torch.equal(red.out[:3 * K].view(K, 3), ra)
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("n", [256, 100, 37])
def test_bit_packed_binary_targets_equal_uint8_targets(gg, cuda_device, n):
    """GG_TARGET_BITS (two-level masks, 1 bit per pixel) gives bit for bit the loss and gradients of the same masks as
    uint8 images, through the planned fused step and through the host-buffer step."""
    import numpy as np
    from gaussigan_b200.host import HostSilhouetteStep
    from gaussigan_b200.plan import SilhouetteLossStep
    B, K = 4, 16
    inp = rp.synth_inputs(B, K, seed=9)
    tgt = rp.synth_target_masks(B, n, seed=n)                         # {-1,+1}
    d = _dev(inp, cuda_device)
    raw = ((tgt.reshape(B, n, n) + 1.0) * 127.5).round().to(torch.uint8)
    packed = torch.from_numpy(np.packbits((tgt.reshape(B, n, n) > 0).numpy(), axis=-1, bitorder="little"))
    assert packed.shape == (B, n, (n + 7) // 8)
    step = SilhouetteLossStep(B, K, n, cuda_device)
    step.bind(d["mus"].reshape(B, K, 3).contiguous(), d["sigma"], *(d[k].reshape(B).contiguous() for k in ("rotx", "roty", "rotz")))
    ref = [t.clone() for t in step.loss_backward(raw.to(cuda_device))] + [step.loss_partials.clone()]
    out = [t.clone() for t in step.loss_backward(packed.contiguous().to(cuda_device), bits=True)] + [step.loss_partials.clone()]
    for a, b in zip(out, ref):
        assert torch.equal(a, b)
    hu = HostSilhouetteStep(B, K, n, cuda_device, depth=2)
    hb = HostSilhouetteStep(B, K, n, cuda_device, depth=2, target_bits=True)
    ru = hu.submit(hu.make_host_inputs(inp, tgt)).wait()
    rb = hb.submit(hb.make_host_inputs(inp, tgt)).wait()
    torch.cuda.synchronize()
    assert hb.h2d_bytes < hu.h2d_bytes and ru.loss == rb.loss
    assert torch.equal(ru.dmus, rb.dmus) and torch.equal(ru.dsigma, rb.dsigma) and torch.equal(ru.drot, rb.drot)


def test_geometry_kernels_vs_reference_source_golden(gg, cuda_device):
    """sigma_from_frame, the 3x3 SVD kernel and deform (CUDA) against Model.get_musigma / Model.transform_mu_sigma
    executed from the reference source (tests/golden/geometry_k8.npz), outputs and gradients w.r.t. the head
    pre-activations."""
    g = load("geometry_k8")
    K = g["cannmu"].shape[1]
    B = g["mus"].shape[0]
    lc = [torch.from_numpy(g[f"pre_c{i}"]).to(cuda_device).requires_grad_(True) for i in range(4)]
    ld = [torch.from_numpy(g[f"pre_d{i}"]).to(cuda_device).requires_grad_(True) for i in range(6)]
    cmu = torch.tanh(lc[0]).reshape(1, K, 1, 3)
    csig = gg.sigma_from_frame(torch.tanh(lc[1]).reshape(1, K, 3), torch.tanh(lc[2]).reshape(1, K, 3), lc[3].reshape(1, K, 3))
    t = [torch.tanh(p) for p in ld]
    mus, sigmas, theta = gg.deform(t[0], t[1], t[2], t[3], t[4], t[5], cmu, csig)
    # float32 construction (:400) on top of tanh evaluated by torch on the GPU (1-2 ulp from the CPU's): ulp-level freedom
    assert_close_norm(csig, g["cannsigma"], 1e-6, "cannsigma")
    assert_close_norm(mus, g["mus"], 1e-6, "mus")
    assert_close_norm(theta, g["theta"], 1e-6, "theta")
    assert_close_norm(sigmas, g["sigmas"], 5e-6, "sigmas")
    w = weights_like([mus.shape, sigmas.shape, theta.shape], g["wseed"])
    ((mus * w[0].to(cuda_device)).sum() + (sigmas * w[1].double().to(cuda_device)).sum() + (theta * w[2].to(cuda_device)).sum()).backward()
    for i, p in enumerate(lc):
        assert_close_norm(p.grad, g[f"grad_c{i}"], GRAD_RTOL, f"grad_c{i}")
    for i, p in enumerate(ld):
        assert_close_norm(p.grad, g[f"grad_d{i}"], GRAD_RTOL, f"grad_d{i}")
