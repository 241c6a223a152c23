import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import gaussigan_b200 as gg
from oracle import reference_port as rp
dev = torch.device('cuda:0')
def todev(inp): return {k: v.to(dev) for k, v in inp.items()}
def rel(a, b): return ((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()
for (B, K, sizes) in [(4, 8, (32, 64, 128, 256)), (3, 16, (256,)), (2, 6, (32, 100)), (2, 12, (64,)), (1, 1, (8,)), (2, 5, (33, 7))]:
    inp = rp.synth_inputs(B, K, seed=56, angle=30.)
    d = todev(inp)
    ref = rp.get_landmarks(inp['mus'], inp['sigma'], inp['rotx'], inp['roty'], inp['rotz'], 0, 0, -2., sizes=sizes)
    out = gg.get_landmarks(d['mus'], d['sigma'], d['rotx'], d['roty'], d['rotz'], 0, 0, -2., sizes=sizes)
    outc = gg.get_landmarks(d['mus'], d['sigma'], d['rotx'], d['roty'], d['rotz'], 0, 0, -2., sizes=sizes, layout='nchw')
    for r, o, oc in zip(ref, out, outc):
        print(B, K, tuple(r.shape), 'nhwc maxabs', (o.cpu() - r).abs().max().item(), 'nchw', (oc.permute(0, 2, 3, 1).cpu() - r).abs().max().item())
    m = gg.render_silhouette(d['mus'], d['sigma'], d['rotx'], d['roty'], d['rotz'], 0, 0, -2., size=sizes[-1])
    print('   mask maxabs', (m.cpu() - ref[-1].sum(-1)).abs().max().item(), 'max', ref[-1].sum(-1).max().item())
    mu2d, s2d = gg.project(d['mus'], d['sigma'], d['rotx'], d['roty'], d['rotz'], 0, 0, -2.)
    rmu, rs = rp.project(inp['mus'], inp['sigma'], inp['rotx'], inp['roty'], inp['rotz'], 0, 0, -2.)
    print('   proj', (mu2d.cpu() - rmu).abs().max().item(), rel(s2d, rs))
# gradient check, mask mode
B, K, n = 4, 16, 256
inp = rp.synth_inputs(B, K, seed=56); tgt = rp.synth_target_masks(B, n)
mask_ref, loss_ref, dmu_ref, dsig_ref, drot_ref = rp.render_loss_and_grads(inp, tgt, n)
d = todev(inp)
mus = d['mus'].clone().requires_grad_(True); sig = d['sigma'].clone().requires_grad_(True); roty = d['roty'].clone().requires_grad_(True)
m = gg.render_silhouette(mus, sig, d['rotx'], roty, d['rotz'], 0, 0, -2., size=n)
loss = ((tgt.to(dev)[..., 0] + 1) * 0.5 - m).abs().mean(); loss.backward()
print('loss', loss.item(), loss_ref.item())
print('grad mask-mode: dmus', rel(mus.grad, dmu_ref), 'dsigma raw', rel(sig.grad, dsig_ref), 'droty', rel(roty.grad, drot_ref))
# gradient check, maps mode, random upstream, all inputs incl rotx/rotz nonzero
for layout in ('nhwc', 'nchw'):
  for (B, K, sizes) in [(3, 16, (32, 64, 128, 256)), (2, 6, (32, 100)), (2, 12, (64,)), (2, 5, (33, 7)), (2, 80, (64,))]:
    inp = rp.synth_inputs(B, K, seed=5, angle=30.)
    inp['rotx'] = torch.randn(B, 1) * 10; inp['rotz'] = torch.randn(B, 1) * 10
    ts = {k: v.clone().requires_grad_(True) for k, v in inp.items()}
    ref = rp.get_landmarks(ts['mus'], ts['sigma'], ts['rotx'], ts['roty'], ts['rotz'], 0.1, -0.2, -2.5, 1.2, sizes=sizes)
    g = torch.Generator().manual_seed(3); gb = [torch.randn(r.shape, generator=g) for r in ref]
    sum((r * q).sum() for r, q in zip(ref, gb)).backward()
    dd = {k: v.to(dev).clone().requires_grad_(True) for k, v in inp.items()}
    out = gg.get_landmarks(dd['mus'], dd['sigma'], dd['rotx'], dd['roty'], dd['rotz'], 0.1, -0.2, -2.5, 1.2, sizes=sizes, layout=layout)
    if layout == 'nchw': out = [o.permute(0, 2, 3, 1) for o in out]
    sum((o * q.to(dev)).sum() for o, q in zip(out, gb)).backward()
    print(layout, B, K, sizes, 'fwd', max((o.cpu() - r).abs().max().item() for o, r in zip(out, ref)), 'grads', {k: rel(dd[k].grad, ts[k].grad) for k in ts})
# get_gaussian_maps_2d + infer
inp = rp.synth_inputs(2, 8, seed=9); d = todev(inp)
a = gg.get_landmarks_infer(d['mus'], d['sigma'], d['rotx'], d['roty'], d['rotz'], 0., 0., -2.)
b = rp.get_landmarks_infer(inp['mus'], inp['sigma'], inp['rotx'], inp['roty'], inp['rotz'], 0., 0., -2.)
print('infer', [(x.cpu() - y).abs().max().item() for x, y in zip(a, b)])
mu2 = b[0].double().clone().requires_grad_(True); s2 = b[1].double().clone().requires_grad_(True)
r = rp.get_gaussian_maps_2d(mu2, s2, [64, 64]); g = torch.Generator().manual_seed(3); q = torch.randn(r.shape, generator=g, dtype=torch.float64); (r * q).sum().backward()
mu2g = b[0].double().to(dev).requires_grad_(True); s2g = b[1].double().to(dev).requires_grad_(True)
o = gg.get_gaussian_maps_2d(mu2g, s2g, [64, 64]); (o * q.to(dev)).sum().backward()
print('raster2d fwd', (o.cpu() - r.float()).abs().max().item(), 'dmu', rel(mu2g.grad, mu2.grad), 'dsig', rel(s2g.grad, s2.grad))
# timing
B, K, n = 64, 16, 256
inp = rp.synth_inputs(B, K, seed=56); d = todev(inp)
mus = d['mus'].clone().requires_grad_(True); sig = d['sigma'].clone().requires_grad_(True); roty = d['roty'].clone().requires_grad_(True)
gm = torch.randn(B, n, n, device=dev)
for it in range(3):
    m = gg.render_silhouette(mus, sig, d['rotx'], roty, d['rotz'], 0, 0, -2., size=n); m.backward(gm)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(20):
    m = gg.render_silhouette(mus, sig, d['rotx'], roty, d['rotz'], 0, 0, -2., size=n); m.backward(gm)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print('fwd+bwd mask B64 K16 256: %.3f ms/iter -> %.0f sil/s' % (ms, B / ms * 1e3))
