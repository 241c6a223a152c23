"""Concat-direct vs the dense round trip at the C2/C4 shape (development tool; the bench line carries the same key)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200 import synth
from gaussigan_b200._lib import lib, check, int_array, ptr_array, LAYOUT_NHWC, GG_PARAM_STRIDE, GG_MOMENTS
dev = torch.device("cuda:0")
B, K, n, CX = int(os.environ.get("B", 64)), int(os.environ.get("K", 16)), 256, int(os.environ.get("CX", 64))
C = CX + K
inp = synth.synth_inputs(B, K, seed=56, device=dev)
mus = inp["mus"].reshape(B, K, 3).contiguous().to(dev); sig = inp["sigma"].to(dev)
r = [inp[k].reshape(B).contiguous().to(dev) for k in ("rotx", "roty", "rotz")]
st = lambda: torch.cuda.current_stream(dev).cuda_stream
params = torch.empty(B, K, GG_PARAM_STRIDE, device=dev)
check(lib.gg_project_fwd(mus.data_ptr(), sig.data_ptr(), 1, r[0].data_ptr(), r[1].data_ptr(), r[2].data_ptr(), 0., 0., -2., 1., B, K, None, None, params.data_ptr(), 0, st()), "p")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
x = torch.randn(B, n, n, CX, device=dev)
bufs = [torch.empty(B, n, n, C, device=dev) for _ in range(2)]
gbufs = [torch.randn(B, n, n, C, device=dev) for _ in range(2)]
dense = [torch.empty(B, n, n, K, device=dev) for _ in range(2)]
sz, one, zero = int_array([n]), int_array([1]), int_array([0])
cs, co = int_array([C]), int_array([CX])
i = [0]
def fwd_concat():
    i[0] ^= 1
    check(lib.gg_splat_fwd_concat(params.data_ptr(), B, K, 1, sz, ptr_array([bufs[i[0]].data_ptr()]), cs, co, None, 0, st()), "fc")
def fwd_dense_then_cat():
    i[0] ^= 1
    check(lib.gg_splat_fwd(params.data_ptr(), B, K, 1, sz, ptr_array([dense[i[0]].data_ptr()]), ptr_array([None]), LAYOUT_NHWC, 0, st()), "fd")
    bufs[i[0]][..., CX:] = dense[i[0]]                       # the copy torch.cat performs for the maps' channels
T = check(lib.gg_splat_bwd_tiles_concat(K, 1, sz, one, zero, cs, co), "t")
part = torch.empty(B, T, K, GG_MOMENTS, device=dev)
Td = check(lib.gg_splat_bwd_tiles(K, 1, sz, one, zero, LAYOUT_NHWC), "t")
partd = torch.empty(B, Td, K, GG_MOMENTS, device=dev)
def bwd_concat():
    i[0] ^= 1
    check(lib.gg_splat_bwd_concat(params.data_ptr(), B, K, 1, sz, ptr_array([gbufs[i[0]].data_ptr()]), cs, co, None, part.data_ptr(), T, 0, st()), "bc")
def bwd_slice_then_dense():
    i[0] ^= 1
    gd = gbufs[i[0]][..., CX:].contiguous()                # what autograd's slice backward hands a dense kernel
    check(lib.gg_splat_bwd(params.data_ptr(), B, K, 1, sz, ptr_array([gd.data_ptr()]), ptr_array([None]), LAYOUT_NHWC, partd.data_ptr(), Td, 0, st()), "bd")
mb = B * n * n * K * 4 / 1e6
print(f"B={B} K={K} n={n} concat buffer channels {C} (maps at {CX}): maps = {mb:.0f} MB")
for name, fn in (("forward  concat-direct", fwd_concat), ("forward  dense + copy into the buffer", fwd_dense_then_cat),
                 ("backward concat-direct", bwd_concat), ("backward slice copy + dense kernel", bwd_slice_then_dense)):
    us = timeit(fn)
    print(f"  {name:40s} {us:8.1f} us   ({mb / us * 1e3 / 1e3:.2f} TB/s of maps)")
