"""Per-kernel CUDA-event timing of the C-ABI launches (development tool; not a bench line)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.plan import SilhouetteStep
from gaussigan_b200._lib import lib, check, int_array, ptr_array, LAYOUT_NHWC, LAYOUT_NCHW
from gaussigan_b200 import synth as rp

ap = argparse.ArgumentParser()
ap.add_argument('--B', type=int, default=64); ap.add_argument('--K', type=int, default=16); ap.add_argument('--n', type=int, default=256)
ap.add_argument('--reps', type=int, default=50)
a = ap.parse_args()
dev = torch.device('cuda:0')
B, K, n = a.B, a.K, a.n
inp = rp.synth_inputs(min(B, 256), K, seed=56)
rep = (B + 255) // 256
mus = inp['mus'].reshape(-1, K, 3).repeat(rep, 1, 1)[:B].contiguous().to(dev)
sig = inp['sigma'].repeat(rep, 1, 1, 1)[:B].contiguous().to(dev)
rx = inp['rotx'].reshape(-1).repeat(rep)[:B].contiguous().to(dev); ry = inp['roty'].reshape(-1).repeat(rep)[:B].contiguous().to(dev); rz = rx.clone()
step = SilhouetteStep(B, K, n, dev)
step.bind(mus, sig, rx, ry, rz)
gm = torch.randn(B, n, n, device=dev) / (B * n * n)

def timeit(fn, reps=a.reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3   # us

st = lambda: torch.cuda.current_stream(dev).cuda_stream
pxk = B * n * n * K
def proj_f(): check(lib.gg_project_fwd(mus.data_ptr(), sig.data_ptr(), 1, rx.data_ptr(), ry.data_ptr(), rz.data_ptr(), 0., 0., -2., 1., B, K, None, None, step.params.data_ptr(), 0, st()), 'pf')
def splat_f(): check(lib.gg_splat_fwd(step.params.data_ptr(), B, K, 1, step._sizes, step._null_ptrs, step._mask_ptrs, LAYOUT_NHWC, 0, st()), 'sf')
gmp = ptr_array([gm.data_ptr()])
def splat_b(): check(lib.gg_splat_bwd(step.params.data_ptr(), B, K, 1, step._sizes, step._null_ptrs, gmp, LAYOUT_NHWC, step.partials.data_ptr(), step.T, 0, st()), 'sb')
def proj_b(): check(lib.gg_project_bwd(mus.data_ptr(), sig.data_ptr(), 1, rx.data_ptr(), ry.data_ptr(), rz.data_ptr(), 0., 0., -2., 1., B, K, step.partials.data_ptr(), step.T, None, None, step.dmus.data_ptr(), step.dsigma.data_ptr(), step.drot.data_ptr(), 0, st()), 'pb')
res = {}
for name, fn in [('project_fwd', proj_f), ('splat_fwd_mask', splat_f), ('splat_bwd_mask', splat_b), ('project_bwd', proj_b)]:
    res[name] = timeit(fn)
def both(): step.forward(); step.backward(gm)
res['step_4launches'] = timeit(both)
step.capture(gm)
res['step_graph'] = timeit(lambda: step.graph.replay())
peak_slots = 148 * 128 * 1.965e9
print(f"B={B} K={K} n={n} T={step.T} RPT_FWD={os.environ.get('GG_RPT_FWD','2')} RPT_BWD={os.environ.get('GG_RPT_BWD','2')}")
for k, v in res.items(): print(f"  {k:18s} {v:9.2f} us")
print(f"  splat_fwd: {pxk/res['splat_fwd_mask']*1e6/1e12:.3f} Tpxk/s = {9*pxk/res['splat_fwd_mask']*1e6/peak_slots*100:.1f}% of 9-slot issue roofline")
print(f"  splat_bwd: {pxk/res['splat_bwd_mask']*1e6/1e12:.3f} Tpxk/s = {17*pxk/res['splat_bwd_mask']*1e6/peak_slots*100:.1f}% of 17-slot issue roofline")
print(f"  step(graph): {B/res['step_graph']*1e6:.0f} sil/s = {26*pxk/res['step_graph']*1e6/peak_slots*100:.1f}% of 26-slot roofline")

# maps mode (NHWC, pyramid) fwd+bwd
sizes = (32, 64, 128, 256)
for layout, lname in ((LAYOUT_NHWC, 'nhwc'), (LAYOUT_NCHW, 'nchw')):
    maps = [torch.empty((B, s, s, K), device=dev) for s in sizes]
    gmaps = [torch.randn((B, s, s, K), device=dev) for s in sizes]
    sz = int_array(sizes); mp = ptr_array([m.data_ptr() for m in maps]); gp = ptr_array([m.data_ptr() for m in gmaps]); nul = ptr_array([None]*4)
    T = check(lib.gg_splat_bwd_tiles(K, 4, sz, int_array([1]*4), int_array([0]*4), layout), 't')
    part = torch.empty((B, T, K, 5), device=dev)
    tf = timeit(lambda: check(lib.gg_splat_fwd(step.params.data_ptr(), B, K, 4, sz, mp, nul, layout, 0, st()), 'x'), 20)
    tb = timeit(lambda: check(lib.gg_splat_bwd(step.params.data_ptr(), B, K, 4, sz, gp, nul, layout, part.data_ptr(), T, 0, st()), 'x'), 20)
    byts = sum(B * s * s * K * 4 for s in sizes)
    print(f"  maps {lname} pyramid: fwd {tf:.1f} us = {byts/tf*1e6/1e9:.0f} GB/s ; bwd {tb:.1f} us = {byts/tb*1e6/1e9:.0f} GB/s (T={T})")
    del maps, gmaps, part
