"""Cost of one more kernel in a programmatic-dependent-launch chain inside a CUDA graph (development tool): chains of
the library's smallest kernel (gg_conic_fwd on 16 Gaussians), with GG_PDL on and off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200._lib import lib, check
dev = torch.device("cuda:0")
mu = torch.zeros(1, 16, 2, dtype=torch.float64, device=dev)
sg = torch.eye(2, dtype=torch.float64, device=dev).repeat(1, 16, 1, 1).contiguous()
pr = torch.empty(1, 16, 8, device=dev)
def chain(n):
    st = torch.cuda.current_stream(dev).cuda_stream
    for _ in range(n):
        check(lib.gg_conic_fwd(mu.data_ptr(), sg.data_ptr(), 1, 1, 16, pr.data_ptr(), 0, st), "c")
for pdl in (1, 0):
    lib.gg_set_pdl(pdl)
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side): chain(8)
    torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): chain(256)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(50): g.replay()
    b.record(); torch.cuda.synchronize()
    print(f"pdl={pdl}: {a.elapsed_time(b) / (50 * 256) * 1e3:.2f} us per dependent launch of a 16-thread kernel")
