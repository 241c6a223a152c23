"""One line per BASELINE.json config: device time of the path on one B200 (CUDA events, 8 rotating input sets where the
working set allows), for DESIGN.md.  Parity of the same shapes is covered by tests/test_parity_gpu.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gaussigan_b200 as gg
from gaussigan_b200.plan import SilhouetteStep, SilhouetteLossStep
from gaussigan_b200._lib import lib, check, int_array, ptr_array, LAYOUT_NHWC, GG_PARAM_STRIDE
from gaussigan_b200 import synth as rp
dev = torch.device("cuda:0")

def timeit(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

def inputs(B, K):
    base = rp.synth_inputs(min(B, 256), K, seed=56)
    rep = (B + 255) // 256
    f = lambda t: t.repeat(rep, *([1] * (t.dim() - 1)))[:B].contiguous().to(dev)
    return f(base["mus"].reshape(-1, K, 3)), f(base["sigma"]), f(base["rotx"].reshape(-1)), f(base["roty"].reshape(-1)), f(base["rotz"].reshape(-1))

def mask_step(B, K, n, reps):
    mus, sig, rx, ry, rz = inputs(B, K)
    st = SilhouetteStep(B, K, n, dev); st.bind(mus, sig, rx, ry, rz)
    gm = torch.randn(B, n, n, device=dev) / (B * n * n)
    st.capture(gm)
    return timeit(lambda: st.graph.replay(), reps)

def maps_fwd(B, K, sizes, reps):
    mus, sig, rx, ry, rz = inputs(B, K)
    params = torch.empty(B, K, GG_PARAM_STRIDE, device=dev)
    maps = [torch.empty(B, s, s, K, device=dev) for s in sizes]
    sz = int_array(sizes); mp = ptr_array([m.data_ptr() for m in maps]); nul = ptr_array([None] * len(sizes))
    stream = lambda: torch.cuda.current_stream(dev).cuda_stream
    def fn():
        check(lib.gg_project_fwd(mus.data_ptr(), sig.data_ptr(), 1, rx.data_ptr(), ry.data_ptr(), rz.data_ptr(), 0., 0., -2., 1., B, K, None, None, params.data_ptr(), 0, stream()), "p")
        check(lib.gg_splat_fwd(params.data_ptr(), B, K, len(sizes), sz, mp, nul, LAYOUT_NHWC, 0, stream()), "s")
    us = timeit(fn, reps)
    return us, sum(m.numel() * 4 for m in maps)

rows = []
us, byts = maps_fwd(16, 8, (128,), 200)
rows.append(("C1 maps fwd, B16 K8 128^2", us, 16, f"{byts / us * 1e6 / 1e9:.0f} GB/s written (launch-bound: 2 kernels)"))
us = mask_step(64, 16, 256, 200)
rows.append(("C2 mask fwd+bwd, B64 K16 256^2 (same buffers, L2-warm)", us, 64, ""))
us = mask_step(256, 16, 256, 50)
rows.append(("C3 shape: mask fwd+bwd, B256 K16 256^2 (render part)", us, 256, "whole dynamic step: tools/time_c3.py"))
us, byts = maps_fwd(128, 16, (32, 64, 128, 256), 50)
rows.append(("C4 maps pyramid fwd, B128 K16", us, 128, f"{byts / us * 1e6 / 1e9:.0f} GB/s written"))
us = mask_step(512, 64, 512, 5)
rows.append(("C5 per-GPU shard: mask fwd+bwd, B512 K64 512^2", us, 512, "x8 GPUs = the B4096 sweep"))
for name, us, B, note in rows:
    print(f"| {name} | {us:.1f} us | {B / us * 1e6:,.0f} silhouettes/s | {note} |")
