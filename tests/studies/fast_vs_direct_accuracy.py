"""Accuracy of the forward-differenced kernels against the direct ones on the GPU (development study): seeded sweeps
over coarse (8-24 px) and reference-sized (32-64 px) grids, Gaussians from far sub-pixel to larger than the frame; prints
the worst gradient errors of both arithmetic modes against the float64 oracle.  Found the re-centring cancellation of
guarded Gaussians.  Run on a GPU box: python tests/studies/fast_vs_direct_accuracy.py"""
import sys, torch
import os; ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import gaussigan_b200 as gg
from oracle import reference_port as rp
from util import max_rel
dev = torch.device('cuda:0')
for name, sizes_pool in (("coarse", [8, 16, 24]), ("ref", [32, 40, 48, 64])):
    rng = torch.Generator().manual_seed(123)
    worst = {True: [0, 0, 0], False: [0, 0, 0]}; wf = {True: 0, False: 0}
    cnt = 0
    for case in range(400):
        B, K = 2, int(torch.randint(1, 25, (1,), generator=rng))
        n = sizes_pool[int(torch.randint(0, len(sizes_pool), (1,), generator=rng))]
        mus = ((torch.rand(B, K, 1, 3, generator=rng) - 0.5) * torch.tensor([3.0, 3.0, 1.0])).float()
        q, _ = torch.linalg.qr(torch.randn(B, K, 3, 3, generator=rng, dtype=torch.float64))
        lo, hi = torch.log(torch.tensor(3e-4, dtype=torch.float64)), torch.log(torch.tensor(0.6, dtype=torch.float64))
        var = torch.exp(torch.rand(B, K, 3, generator=rng, dtype=torch.float64) * (hi - lo) + lo)
        sig = (q * var[..., None, :]) @ q.transpose(-1, -2)
        rot = [(torch.rand(B, 1, generator=rng) - 0.5) * a for a in (40.0, 360.0, 40.0)]
        with_mask = bool(torch.randint(0, 2, (1,), generator=rng))
        lv = [x.clone().requires_grad_(True) for x in (mus, sig, rot[1])]
        ref = rp.get_landmarks(lv[0], lv[1], rot[0], lv[2], rot[2], 0, 0, -2.5, sizes=(n,))[0]
        skip = (not torch.isfinite(ref).all() or float(ref.detach().abs().max()) < 1e-6 or float(ref.detach().max()) > 1.5)
        w = torch.randn(B, n, n, K, generator=rng); wm = torch.randn(B, n, n, 1, generator=rng)
        if skip: continue
        cnt += 1
        ((ref * w).sum() + ((ref.sum(-1, keepdim=True) * wm).sum() if with_mask else 0.0)).backward()
        for fast in (True, False):
            gg.set_fast_mask(fast)
            dv = [x.to(dev).requires_grad_(True) for x in (mus, sig, rot[1])]
            out, mask = gg.get_landmarks_with_mask(dv[0], dv[1], rot[0].to(dev), dv[2], rot[2].to(dev), 0, 0, -2.5, sizes=(n,))
            wf[fast] = max(wf[fast], max_rel(mask, ref.detach().sum(-1, keepdim=True)))
            ((out[0] * w.to(dev)).sum() + ((mask * wm.to(dev)).sum() if with_mask else 0.0)).backward()
            e = [max_rel(a.grad, b.grad) if float(b.grad.abs().max()) > 1e-12 else 0.0 for a, b in zip(dv, lv)]
            worst[fast] = [max(a, b) for a, b in zip(worst[fast], e)]
    print(name, cnt, 'cases; worst (mus, sigma, roty) fast', ['%.1e' % x for x in worst[True]], 'direct', ['%.1e' % x for x in worst[False]], 'mask fwd fast %.1e direct %.1e' % (wf[True], wf[False]))
