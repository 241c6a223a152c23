"""CUDA-event timing of the NHWC maps backward and forward (gg_splat_bwd with upstream gradients of the K-channel maps),
single scale and pyramid (development tool; not a bench line)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200._lib import lib, check, int_array, ptr_array, LAYOUT_NHWC
from gaussigan_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument('--B', type=int, default=64); ap.add_argument('--K', type=int, default=16)
ap.add_argument('--sizes', type=str, default='256'); ap.add_argument('--reps', type=int, default=20)
ap.add_argument('--sets', type=int, default=2, help='rotating gradient sets (> L2 in total)')
ap.add_argument('--graph', type=int, default=0, help='time a CUDA graph of this many back-to-back launches instead of eager launches')
a = ap.parse_args()
dev = torch.device('cuda:0')
B, K = a.B, a.K
sizes = tuple(int(s) for s in a.sizes.split(','))
inp = synth.synth_inputs(min(B, 256), K, seed=56)
rep = (B + 255) // 256
mus = inp['mus'].reshape(-1, K, 3).repeat(rep, 1, 1)[:B].contiguous().to(dev)
sig = inp['sigma'].repeat(rep, 1, 1, 1)[:B].contiguous().to(dev)
rx = inp['rotx'].reshape(-1).repeat(rep)[:B].contiguous().to(dev); ry = inp['roty'].reshape(-1).repeat(rep)[:B].contiguous().to(dev)
params = torch.empty(B, K, 8, device=dev)
st = lambda: torch.cuda.current_stream(dev).cuda_stream
check(lib.gg_project_fwd(mus.data_ptr(), sig.data_ptr(), 1, rx.data_ptr(), ry.data_ptr(), rx.data_ptr(), 0., 0., -2., 1., B, K, None, None, params.data_ptr(), 0, st()), 'pf')
ns = len(sizes)
sz = int_array(sizes); nul = ptr_array([None] * ns)
T = check(lib.gg_splat_bwd_tiles(K, ns, sz, int_array([1] * ns), int_array([0] * ns), LAYOUT_NHWC), 't')
part = torch.empty((B, T, K, 5), device=dev)
gsets = [[torch.randn((B, s, s, K), device=dev) for s in sizes] for _ in range(a.sets)]
gps = [ptr_array([m.data_ptr() for m in g]) for g in gsets]
byts = sum(B * s * s * K * 4 for s in sizes)
def run(i): check(lib.gg_splat_bwd(params.data_ptr(), B, K, ns, sz, gps[i % a.sets], nul, LAYOUT_NHWC, part.data_ptr(), T, 0, st()), 'x')
for i in range(3): run(i)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
def timed(fn):
    if a.graph:
        side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for i in range(a.graph): fn(i)
        torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(a.graph): fn(i)
        for _ in range(3): g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(a.reps): g.replay()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (a.reps * a.graph) * 1e3
    e0.record()
    for i in range(a.reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.reps * 1e3
us = timed(run)
print(f"B={B} K={K} sizes={sizes} T={T}: bwd {us:.1f} us = {byts/us*1e6/1e9:.0f} GB/s ({byts/1e6:.0f} MB per launch)")
# forward into the same buffers
def runf(i): check(lib.gg_splat_fwd(params.data_ptr(), B, K, ns, sz, gps[i % a.sets], nul, LAYOUT_NHWC, 0, st()), 'f')
for i in range(3): runf(i)
torch.cuda.synchronize()
us = timed(runf)
print(f"B={B} K={K} sizes={sizes}: fwd {us:.1f} us = {byts/us*1e6/1e9:.0f} GB/s")
