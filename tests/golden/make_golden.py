"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE (build container only).

The reference functions are pulled out of /root/reference with ``ast`` and run against the torch-backed
TensorFlow shim (oracle/tf_shim.py); nothing of the reference is copied into the repository.  Outputs
(forward maps, and gradients obtained by autograd through the reference's code) are stored as small
fixtures so that the GPU box -- where /root/reference does not exist -- can still check against them.

    python tests/golden/make_golden.py          # rewrites the .npz files next to this script

Large maps are stored sub-sampled (every 8th pixel of the 128 and 256 scales) plus their per-sample
sums, to keep the fixtures small.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_port as rp  # noqa: E402
from oracle import tf_shim  # noqa: E402

REF = "/root/reference"
SUB = 8


def _weights(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g, dtype=torch.float32) for s in shapes]


def _pack_maps(prefix, maps, out):
    for m in maps:
        n = m.shape[1]
        a = m.detach().numpy()
        out[f"{prefix}{n}_sum"] = a.astype(np.float64).sum(axis=(1, 2))
        out[f"{prefix}{n}"] = a if n <= 64 else a[:, ::SUB, ::SUB, :]


def train_case(name, B, K, seed, angle, rxz_scale, t, focal):
    fns = tf_shim.load_reference_functions(os.path.join(REF, "mask/mask_gen_dynamic.py"),
                                           ["get_landmarks", "get_gaussian_maps_2d"], K)
    inp = rp.synth_inputs(B, K, seed=seed, angle=angle)
    if rxz_scale:
        g = torch.Generator().manual_seed(seed + 1)
        inp["rotx"] = (torch.randn(B, 1, generator=g) * rxz_scale).float()
        inp["rotz"] = (torch.randn(B, 1, generator=g) * rxz_scale).float()
    leaves = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    maps = fns["get_landmarks"](leaves["mus"], leaves["sigma"], leaves["rotx"], leaves["roty"], leaves["rotz"],
                                t[0], t[1], t[2], focal)
    assert [m.shape[1] for m in maps] == [32, 64, 128, 256] and all(m.dtype == torch.float32 for m in maps)
    w = _weights([m.shape for m in maps], seed + 2)
    sum((m * q).sum() for m, q in zip(maps, w)).backward()
    out = {f"in_{k}": v.numpy() for k, v in inp.items()}
    out.update(t=np.asarray(t, np.float64), focal=np.float64(focal), wseed=np.int64(seed + 2))
    _pack_maps("maps", maps, out)
    for k, v in leaves.items():
        out[f"grad_{k}"] = v.grad.numpy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


def infer_case(name, B, K, seed):
    fns = tf_shim.load_reference_functions(os.path.join(REF, "mask/infer_masks.py"),
                                           ["get_landmarks", "get_gaussian_maps_2d"], K)
    inp = rp.synth_inputs(B, K, seed=seed)
    mu2d, sigma2d, gm = fns["get_landmarks"](inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"],
                                             0., 0., -2.)
    assert mu2d.dtype == torch.float32 and gm.dtype == torch.float32 and gm.shape[1] == 256
    out = {f"in_{k}": v.numpy() for k, v in inp.items()}
    out.update(mu2d=mu2d.numpy(), sigma2d=sigma2d.numpy())
    _pack_maps("maps", [gm], out)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def raster_case(name, B, K, n, seed):
    """get_gaussian_maps_2d alone, with float64 inputs and gradients (the frozen-graph boundary)."""
    fns = tf_shim.load_reference_functions(os.path.join(REF, "mask/mask_gen_dynamic.py"), ["get_gaussian_maps_2d"], K)
    inp = rp.synth_inputs(B, K, seed=seed)
    mu2d, sigma2d = rp.project(inp["mus"], inp["sigma"], inp["rotx"], inp["roty"], inp["rotz"], 0., 0., -2.)
    mu = mu2d.clone().requires_grad_(True)
    sg = sigma2d.clone().requires_grad_(True)
    g = fns["get_gaussian_maps_2d"](mu, sg, [n, n])
    assert g.dtype == torch.float64 and tuple(g.shape) == (B, n, n, K)
    w = _weights([g.shape], seed + 2)[0].double()
    (g * w).sum().backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), mu2d=mu2d.numpy(), sigma2d=sigma2d.numpy(),
                        maps=g.detach().numpy().astype(np.float32), grad_mu2d=mu.grad.numpy(),
                        grad_sigma2d=sg.grad.numpy(), wseed=np.int64(seed + 2), n=np.int64(n))
    print(name, tuple(g.shape))


def geometry_case(name, B, K, seed):
    """Rows a6 / a7: Model.get_musigma and Model.transform_mu_sigma executed as they stand (oracle/tf_shim.py feeds
    pre-drawn head pre-activations through ``tf.layers.dense``); outputs and autograd gradients w.r.t. the
    pre-activations are stored."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    pre_c = [rn(1, K * 3) for _ in range(4)]                                   # mus, v0, v1, u heads (call order)
    pre_d = [rn(B, K * 3), rn(B, K * 3), rn(B, K), rn(B, K), rn(B, K), rn(B, 1)]   # mu_t, Sscale, rx, ry, rz, theta
    lc = [p.clone().requires_grad_(True) for p in pre_c]
    ld = [p.clone().requires_grad_(True) for p in pre_d]
    path = os.path.join(REF, "mask/mask_gen_dynamic.py")
    f1 = tf_shim.load_reference_methods(path, "Model", ["get_musigma"], K, tf_shim.DenseQueue([K * 3], lc))
    cmu, csig = f1["get_musigma"](None)
    f2 = tf_shim.load_reference_methods(path, "Model", ["transform_mu_sigma"], K, tf_shim.DenseQueue([K * 3, K, 1], ld))
    mus, sigmas, theta = f2["transform_mu_sigma"](None, torch.zeros(B, 1, 1, 8), cmu, csig)
    w = _weights([mus.shape, sigmas.shape, theta.shape], seed + 2)
    ((mus * w[0]).sum() + (sigmas * w[1].double()).sum() + (theta * w[2]).sum()).backward()
    out = {f"pre_c{i}": p.numpy() for i, p in enumerate(pre_c)}
    out.update({f"pre_d{i}": p.numpy() for i, p in enumerate(pre_d)})
    out.update(cannmu=cmu.detach().numpy(), cannsigma=csig.detach().numpy(), mus=mus.detach().numpy(),
               sigmas=sigmas.detach().numpy(), theta=theta.detach().numpy(), wseed=np.int64(seed + 2))
    out.update({f"grad_c{i}": p.grad.numpy() for i, p in enumerate(lc)})
    out.update({f"grad_d{i}": p.grad.numpy() for i, p in enumerate(ld)})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference (build container only)")
    torch.manual_seed(0)
    # the authors' own K (mask/train_giraffe.sh: 6, mask/train_manuel.sh: 12); call sites pass rotx=rotz=0, t=(0,0,-2)
    train_case("train_k6", B=2, K=6, seed=56, angle=0.0, rxz_scale=0.0, t=(0.0, 0.0, -2.0), focal=1.0)
    train_case("train_k12_general", B=2, K=12, seed=7, angle=30.0, rxz_scale=12.0, t=(0.1, -0.2, -2.5), focal=1.3)
    train_case("train_k16", B=3, K=16, seed=11, angle=30.0, rxz_scale=0.0, t=(0.0, 0.0, -2.0), focal=1.0)
    infer_case("infer_k6", B=2, K=6, seed=56)
    raster_case("raster_k8", B=2, K=8, n=48, seed=21)
    geometry_case("geometry_k8", B=3, K=8, seed=31)
