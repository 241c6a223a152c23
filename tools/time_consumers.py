"""Device time of the fused consumers at the C2 shape (B=64, K=16, 256^2): cycle L1 (two renders, loss + both gradient
moment sets, no maps), colourise from Gaussians (uint8 frames), turntable (60 angles)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gaussigan_b200 as gg
from gaussigan_b200 import synth
from gaussigan_b200._lib import lib, check, GG_PARAM_STRIDE, GG_MOMENTS
dev = torch.device("cuda:0"); B, K, n = 64, 16, 256
def timeit(fn, reps=100):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
st = lambda: torch.cuda.current_stream(dev).cuda_stream
def params_of(seed):
    inp = synth.synth_inputs(B, K, seed=seed, device=dev)
    p = torch.empty(B, K, GG_PARAM_STRIDE, device=dev)
    mus = inp["mus"].reshape(B, K, 3).contiguous().to(dev); sig = inp["sigma"].to(dev)
    r = [inp[k].reshape(B).contiguous().to(dev) for k in ("rotx", "roty", "rotz")]
    check(lib.gg_project_fwd(mus.data_ptr(), sig.data_ptr(), 1, r[0].data_ptr(), r[1].data_ptr(), r[2].data_ptr(), 0., 0., -2., 1., B, K, None, None, p.data_ptr(), 0, st()), "p")
    return p, inp
pa, inp = params_of(1); pb, _ = params_of(2)
T = lib.gg_cycle_tiles(n); nl = lib.gg_cycle_loss_count(B, K, n)
qa = torch.empty(B, T, K, GG_MOMENTS, device=dev); qb = torch.empty_like(qa); lp = torch.empty(nl, device=dev)
us = timeit(lambda: check(lib.gg_cycle_l1_fused(pa.data_ptr(), pb.data_ptr(), B, K, n, qa.data_ptr(), qb.data_ptr(), T, lp.data_ptr(), 0, st()), "c"))
pxk = B * n * n * K
print(f"cycle L1 fused (2 renders + loss + 2 moment sets): {us:.1f} us = {pxk / us * 1e6 / 1e12:.2f} T pixel-Gaussians/s ({B / us * 1e6:,.0f} silhouette pairs/s); "
      f"the two [B,n,n,K] maps it never writes would be {2 * pxk * 4 / 1e6:.0f} MB")
cols = torch.rand(K, 3, device=dev); frames = torch.empty(B, n, n, 3, dtype=torch.uint8, device=dev)
us = timeit(lambda: check(lib.gg_colorize(pa.data_ptr(), None, cols.data_ptr(), B, K, n, frames.data_ptr(), 2, 0, st()), "z"))
print(f"colourise from Gaussians -> uint8 frames: {us:.1f} us ({B / us * 1e6:,.0f} frames/s)")
mu3d = inp["mus"][:1].to(dev); sig3d = inp["sigma"][:1].to(dev)
import numpy as np
us = timeit(lambda: gg.turntable(mu3d, sig3d, 10.0, np.linspace(0, 360, 60), cols, size=n), 30)
print(f"turntable, 60 angles, projection + colourised uint8 frames (public API, eager): {us:.1f} us per turntable")
