"""CPU tests of the oracle itself: golden vectors (made by executing the reference source), known answers,
the independent closed form, and -- when /root/reference is present -- the reference source live."""
import os

import numpy as np
import pytest
import torch

from oracle import closed_form as cf
from oracle import reference_port as rp
from util import SIZES, assert_close_norm, inputs_of, load, max_rel, sub, weights_like

REF = "/root/reference"


def test_tf_linspace_formula():
    for n in (1, 2, 7, 32, 256, 512):
        x = rp.tf_linspace_f32(n)
        assert x.dtype == np.float32 and x.shape == (n,) and x[0] == np.float32(-1.0)
        if n > 1:
            step = np.float32(np.float32(2.0) / np.float32(n - 1))
            assert x[n // 2] == np.float32(np.float32(-1.0) + np.float32(step * np.float32(n // 2)))
            assert abs(float(x[-1]) - 1.0) <= 2e-7          # last element within 1 ulp of +1 (unpinned in TF)


@pytest.mark.parametrize("name", ["train_k6", "train_k12_general", "train_k16"])
def test_port_matches_reference_source_golden(name):
    g = load(name)
    inp = inputs_of(g)
    leaves = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    t = g["t"]
    maps = rp.get_landmarks(leaves["mus"], leaves["sigma"], leaves["rotx"], leaves["roty"], leaves["rotz"],
                            float(t[0]), float(t[1]), float(t[2]), float(g["focal"]))
    for m in maps:
        n = m.shape[1]
        assert m.dtype == torch.float32
        np.testing.assert_allclose(sub(m.detach()).numpy(), g[f"maps{n}"], rtol=0, atol=1e-7)
        np.testing.assert_allclose(m.detach().double().sum((1, 2)).numpy(), g[f"maps{n}_sum"], rtol=1e-9)
    w = weights_like([m.shape for m in maps], g["wseed"])
    sum((m * q).sum() for m, q in zip(maps, w)).backward()
    for k, v in leaves.items():
        assert_close_norm(v.grad, g[f"grad_{k}"], 1e-6, f"{name} grad_{k}")


def test_port_infer_matches_golden():
    g = load("infer_k6")
    inp = inputs_of(g)
    mu2d, s2d, maps = rp.get_landmarks_infer(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0., 0., -2.)
    assert mu2d.dtype == torch.float32 and maps.dtype == torch.float32
    np.testing.assert_array_equal(mu2d.numpy(), g["mu2d"])
    np.testing.assert_array_equal(s2d.numpy(), g["sigma2d"])
    np.testing.assert_allclose(sub(maps).numpy(), g["maps256"], rtol=0, atol=2e-7)


def test_port_raster_matches_golden():
    g = load("raster_k8")
    mu = torch.from_numpy(g["mu2d"]).requires_grad_(True)
    sg = torch.from_numpy(g["sigma2d"]).requires_grad_(True)
    n = int(g["n"])
    m = rp.get_gaussian_maps_2d(mu, sg, [n, n])
    np.testing.assert_allclose(m.detach().float().numpy(), g["maps"], rtol=0, atol=1e-7)
    w = weights_like([m.shape], g["wseed"])[0].double()
    (m * w).sum().backward()
    assert_close_norm(mu.grad, g["grad_mu2d"], 1e-9, "grad_mu2d")
    assert_close_norm(sg.grad, g["grad_sigma2d"], 1e-9, "grad_sigma2d")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
@pytest.mark.parametrize("K,general", [(6, False), (16, True)])
def test_port_matches_reference_source_live(K, general):
    from oracle import tf_shim
    fns = tf_shim.load_reference_functions(os.path.join(REF, "mask/mask_gen_dynamic.py"),
                                           ["get_landmarks", "get_gaussian_maps_2d"], K)
    inp = rp.synth_inputs(3, K, seed=123, angle=45.0)
    t, f = (0.0, 0.0, -2.0), 1.0
    if general:
        inp["rotx"] = torch.tensor([[5.0], [-17.0], [0.5]])
        inp["rotz"] = torch.tensor([[-3.0], [8.0], [33.0]])
        t, f = (0.2, 0.1, -3.0), 0.8
    a = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    b = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    ra = fns["get_landmarks"](a["mus"], a["sigma"], a["rotx"], a["roty"], a["rotz"], *t, f)
    rb = rp.get_landmarks(b["mus"], b["sigma"], b["rotx"], b["roty"], b["rotz"], *t, f)
    w = weights_like([m.shape for m in ra], 5)
    for x, y in zip(ra, rb):
        assert torch.equal(x, y)                       # same ops in the same order -> bit-identical
    sum((m * q).sum() for m, q in zip(ra, w)).backward()
    sum((m * q).sum() for m, q in zip(rb, w)).backward()
    for k in a:
        assert_close_norm(b[k].grad, a[k].grad, 1e-12, k)


def test_sphere_known_answer():
    mu = torch.zeros(1, 1, 1, 3)
    S = 0.09 * torch.eye(3, dtype=torch.float64).view(1, 1, 3, 3)
    z = torch.zeros(1, 1)
    mu2d, s2d = rp.project(mu, S, z, z, z, 0, 0, -2.0)
    kmu, ksig = cf.sphere_kat(0.09, -2.0)
    np.testing.assert_allclose(mu2d[0, 0].numpy(), kmu, atol=1e-15)
    np.testing.assert_allclose(s2d[0, 0].numpy(), ksig, rtol=1e-13, atol=1e-16)
    g = rp.get_gaussian_maps_2d(mu2d, s2d, [33, 33])      # odd n puts a pixel centre (almost) on the axis
    assert abs(float(g[0, 16, 16, 0]) - 1.0) < 1e-5 and float(g.max()) <= 1.0


def test_yaw_sign_convention():
    S = 0.09 * torch.eye(3, dtype=torch.float64).view(1, 1, 3, 3)
    z = torch.zeros(1, 1)
    mu = torch.tensor([0.5, 0.0, 0.0]).view(1, 1, 1, 3)
    for yaw in (90.0, -90.0):                               # (0.5,0,0) is rotated onto the optical axis
        mu2d, _ = rp.project(mu, S, z, z + yaw, z, 0, 0, -2.0)
        assert mu2d.abs().max() < 1e-7
    m0, _ = rp.project(mu, S, z, z, z, 0, 0, -2.0)
    assert abs(float(m0[0, 0, 0])) > 0.2 and abs(float(m0[0, 0, 1])) < 1e-12


def test_closed_form_equals_port_for_symmetric_sigma():
    inp = rp.synth_inputs(8, 16, seed=56, angle=30.0)
    inp["rotx"] = torch.linspace(-20, 20, 8).view(8, 1)
    inp["rotz"] = torch.linspace(30, -30, 8).view(8, 1)
    sig = rp.sym(inp["sigma"])
    mu2d, s2d = rp.project(inp["mus"], sig, inp["rotx"], inp["roty"], inp["rotz"], 0.1, 0.2, -2.2, 1.1)
    m2, s2, _ = cf.project_closed(inp["mus"][:, :, 0].numpy(), sig.numpy(), inp["rotx"][:, 0].numpy(),
                                  inp["roty"][:, 0].numpy(), inp["rotz"][:, 0].numpy(), t=(0.1, 0.2, -2.2), focal=1.1)
    assert np.abs(m2 - mu2d.numpy()).max() < 1e-13
    assert max_rel(s2, s2d) < 1e-12
    assert np.abs(cf.raster(m2, s2, 40) - rp.get_gaussian_maps_2d(mu2d, s2d, [40, 40]).numpy()).max() < 1e-12


def test_hand_backward_equals_autograd():
    inp = rp.synth_inputs(3, 8, seed=3, angle=20.0)
    inp["rotx"] = torch.tensor([[4.0], [0.0], [-9.0]])
    inp["rotz"] = torch.tensor([[0.0], [6.0], [15.0]])
    sig = rp.sym(inp["sigma"])
    lv = {k: v.clone().requires_grad_(True) for k, v in dict(inp, sigma=sig).items()}
    maps = rp.get_landmarks(lv["mus"], lv["sigma"], lv["rotx"], lv["roty"], lv["rotz"], 0, 0, -2.0, sizes=(32, 48))
    w = weights_like([m.shape for m in maps], 1)
    sum((m * q).sum() for m, q in zip(maps, w)).backward()
    m2, s2, cache = cf.project_closed(inp["mus"][:, :, 0].numpy(), sig.numpy(), inp["rotx"][:, 0].numpy(),
                                      inp["roty"][:, 0].numpy(), inp["rotz"][:, 0].numpy())
    mom = cf.pixel_moments(m2, s2, [q.numpy() for q in w])
    dmu2, ds2 = cf.conic_backward(s2, mom)
    dm, dS, dang = cf.project_backward(cache, dmu2, ds2)
    assert max_rel(dm, lv["mus"].grad[:, :, 0]) < 1e-6          # autograd's dmus is rounded to float32
    assert max_rel(dS, rp.sym(lv["sigma"].grad)) < 1e-12
    for d, k in zip(dang, ("rotx", "roty", "rotz")):
        assert max_rel(d, lv[k].grad[:, 0]) < 1e-6


def test_sigma_gradient_is_asymmetric_in_the_reference():
    """Documents why gradient parity is stated on the raw d Sigma AND on its symmetric part."""
    inp = rp.synth_inputs(2, 4, seed=1)
    s = inp["sigma"].clone().requires_grad_(True)
    m = rp.get_landmarks(inp["mus"], s, inp["rotx"], inp["roty"], inp["rotz"], 0, 0, -2.0, sizes=(32,))[0]
    m.sum().backward()
    anti = (s.grad - rp.sym(s.grad)).abs().max()
    assert anti > 0.1 * rp.sym(s.grad).abs().max()


def test_sigma_construction_and_deformation_shapes():
    g = torch.Generator().manual_seed(0)
    K, B = 6, 3
    v0, v1 = torch.tanh(torch.randn(1, K, 3, generator=g)), torch.tanh(torch.randn(1, K, 3, generator=g))
    S = rp.sigma_from_frame(v0, v1, torch.randn(1, K, 3, generator=g))
    assert S.dtype == torch.float64 and S.shape == (1, K, 3, 3)
    ev = torch.linalg.eigvalsh(rp.sym(S))
    assert ev.min() > 0.0099 and ev.max() < 0.5101           # u = 0.5*sigmoid + 0.01
    cmu = torch.tanh(torch.randn(1, K, 1, 3, generator=g))
    th = lambda *s: torch.tanh(torch.randn(*s, generator=g))
    mus, sig, theta = rp.deform(th(B, K * 3), th(B, K * 3), th(B, K), th(B, K), th(B, K), th(B, 1), cmu, S)
    assert mus.shape == (B, K, 1, 3) and mus.dtype == torch.float32
    assert sig.shape == (B, K, 3, 3) and sig.dtype == torch.float64 and theta.shape == (B, 1)
    assert (mus - cmu).abs().max() <= 0.1 + 1e-6 and theta.abs().max() <= 180.0
    # zero heads leave the canonical Gaussians untouched (scale 1, rotation I)
    z = torch.zeros
    mus0, sig0, _ = rp.deform(z(B, K * 3), z(B, K * 3), z(B, K), z(B, K), z(B, K), z(B, 1), cmu, S)
    assert torch.allclose(sig0, rp.sym(S).expand(B, K, 3, 3), atol=1e-7)


def test_consumers():
    maps = torch.rand(2, 8, 8, 5)
    assert torch.allclose(rp.sum_composite(maps)[..., 0], maps.sum(-1))
    colors = torch.rand(5, 3)
    rgb = rp.colorize_landmark_maps(maps, colors)
    assert rgb.shape == (2, 8, 8, 3)
    assert torch.allclose(rgb[0, 3, 4], (maps[0, 3, 4, :, None] * colors).max(0).values)


# ---------------------------------------------------------------------------- geometry tails (rows a6 / a7)
def _geometry_port(g):
    """The oracle port on the golden's head pre-activations; gradients w.r.t. the pre-activations by autograd."""
    K = g["cannmu"].shape[1]
    lc = [torch.from_numpy(g[f"pre_c{i}"]).clone().requires_grad_(True) for i in range(4)]
    ld = [torch.from_numpy(g[f"pre_d{i}"]).clone().requires_grad_(True) for i in range(6)]
    cmu = torch.tanh(lc[0]).reshape(1, K, 1, 3)                                    # mask/mask_gen_dynamic.py:378-379
    csig = rp.sigma_from_frame(torch.tanh(lc[1]).reshape(1, K, 3), torch.tanh(lc[2]).reshape(1, K, 3), lc[3].reshape(1, K, 3))
    t = [torch.tanh(p) for p in ld]
    mus, sigmas, theta = rp.deform(t[0], t[1], t[2], t[3], t[4], t[5], cmu, csig)
    return lc, ld, cmu, csig, mus, sigmas, theta


def test_geometry_port_matches_reference_source_golden():
    """sigma_from_frame + deform against Model.get_musigma / Model.transform_mu_sigma EXECUTED from the reference
    source (tests/golden/make_golden.py: geometry_case)."""
    g = load("geometry_k8")
    lc, ld, cmu, csig, mus, sigmas, theta = _geometry_port(g)
    assert torch.equal(cmu.detach(), torch.from_numpy(g["cannmu"])) and torch.equal(csig.detach(), torch.from_numpy(g["cannsigma"]))
    assert torch.equal(mus.detach(), torch.from_numpy(g["mus"])) and torch.equal(theta.detach(), torch.from_numpy(g["theta"]))
    assert_close_norm(sigmas, g["sigmas"], 1e-14, "sigmas")
    w = weights_like([mus.shape, sigmas.shape, theta.shape], g["wseed"])
    ((mus * w[0]).sum() + (sigmas * w[1].double()).sum() + (theta * w[2]).sum()).backward()
    for i, p in enumerate(lc):
        assert_close_norm(p.grad, g[f"grad_c{i}"], 1e-6, f"grad_c{i}")          # float32 head chain
    for i, p in enumerate(ld):
        assert_close_norm(p.grad, g[f"grad_d{i}"], 1e-6, f"grad_d{i}")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_geometry_port_matches_reference_source_live():
    from oracle import tf_shim
    K, B = 5, 4
    gen = torch.Generator().manual_seed(77)
    rn = lambda *s: torch.randn(*s, generator=gen)
    pre_c = [rn(1, K * 3) for _ in range(4)]
    pre_d = [rn(B, K * 3), rn(B, K * 3), rn(B, K), rn(B, K), rn(B, K), rn(B, 1)]
    path = os.path.join(REF, "mask/mask_gen_dynamic.py")
    f1 = tf_shim.load_reference_methods(path, "Model", ["get_musigma"], K, tf_shim.DenseQueue([K * 3], [p.clone() for p in pre_c]))
    cmu, csig = f1["get_musigma"](None)
    f2 = tf_shim.load_reference_methods(path, "Model", ["transform_mu_sigma"], K,
                                        tf_shim.DenseQueue([K * 3, K, 1], [p.clone() for p in pre_d]))
    mus, sigmas, theta = f2["transform_mu_sigma"](None, torch.zeros(B, 1, 1, 8), cmu, csig)
    pc = rp.sigma_from_frame(torch.tanh(pre_c[1]).reshape(1, K, 3), torch.tanh(pre_c[2]).reshape(1, K, 3), pre_c[3].reshape(1, K, 3))
    t = [torch.tanh(p) for p in pre_d]
    pm, ps, pt = rp.deform(t[0], t[1], t[2], t[3], t[4], t[5], torch.tanh(pre_c[0]).reshape(1, K, 1, 3), pc)
    assert torch.equal(pc, csig) and torch.equal(pm, mus) and torch.equal(pt, theta)
    assert_close_norm(ps, sigmas, 1e-14, "sigmas")


# ------------------------------------------------- the shim's primitives are its own, and agree with the port's
def test_shim_primitives_are_independent_and_equal_the_port():
    """oracle/tf_shim.py must not borrow cos/sin/matmul/linspace from the port it is used to check:
    its numpy definitions and the port's torch definitions agree bit for bit, values and gradients."""
    import inspect
    from oracle import tf_shim
    src = inspect.getsource(tf_shim)
    assert "from .reference_port import" not in src and "import reference_port" not in src
    for name in ("cos_f32", "sin_f32", "matmul3_f32", "tf_linspace_f32"):
        assert getattr(tf_shim, name).__module__ == "oracle.tf_shim"
    g = torch.Generator().manual_seed(3)
    ang = torch.cat([(torch.rand(20000, generator=g) - 0.5) * 8.0, torch.tensor([0.0, np.pi / 2, -np.pi, 3.14159274])]).float()
    for f_s, f_p in ((tf_shim.cos_f32, rp.cos_f32), (tf_shim.sin_f32, rp.sin_f32)):
        a1, a2 = ang.clone().requires_grad_(True), ang.clone().requires_grad_(True)
        y1, y2 = f_s(a1), f_p(a2)
        assert torch.equal(y1, y2)
        w = torch.randn(ang.shape, generator=g)
        (y1 * w).sum().backward(); (y2 * w).sum().backward()
        assert torch.equal(a1.grad, a2.grad)
    A, Bm = torch.randn(50, 1, 3, 3, generator=g), torch.randn(50, 1, 3, 3, generator=g)
    a1, b1, a2, b2 = (t.clone().requires_grad_(True) for t in (A, Bm, A, Bm))
    y1, y2 = tf_shim.matmul3_f32(a1, b1), rp.matmul3_f32(a2, b2)
    assert torch.equal(y1, y2)
    w = torch.randn(y1.shape, generator=g)
    (y1 * w).sum().backward(); (y2 * w).sum().backward()
    assert torch.equal(a1.grad, a2.grad) and torch.equal(b1.grad, b2.grad)
    for n in (1, 2, 3, 32, 33, 64, 100, 128, 256, 512, 2100):
        assert np.array_equal(tf_shim.tf_linspace_f32(n), rp.tf_linspace_f32(n)), n


def test_float32_trig_is_correctly_rounded():
    """The pinned primitive: cos_f32 / sin_f32 are the correctly rounded float32 values -- checked against
    50-digit mpmath on the reference's angle range (yaw = 180 tanh(.) degrees, :458) and beyond."""
    import mpmath
    mpmath.mp.dps = 50
    g = torch.Generator().manual_seed(8)
    deg = torch.cat([180.0 * torch.tanh(torch.randn(1500, generator=g)), (torch.rand(500, generator=g) - 0.5) * 720.0])
    rad = (deg.float() * torch.tensor(np.pi, dtype=torch.float32) / torch.tensor(180.0)).float()
    c, s = rp.cos_f32(rad).numpy(), rp.sin_f32(rad).numpy()
    for i, x in enumerate(rad.numpy()):
        xm = mpmath.mpf(float(x))
        assert np.float32(float(mpmath.cos(xm))) == c[i], (x, c[i])
        assert np.float32(float(mpmath.sin(xm))) == s[i], (x, s[i])
