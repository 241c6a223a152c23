"""Marginal cost of each kernel of the headline step inside the CUDA-graph chain: the 64-step chain timed with one or
both projection kernels left out (development tool; the shortened chains compute nothing meaningful)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.plan import SilhouetteStep
from gaussigan_b200._lib import lib, check, ptr_array, LAYOUT_NHWC
from gaussigan_b200 import synth
dev = torch.device("cuda:0"); B, K, n, R = 64, 16, 256, 8
sets = []
for r in range(R):
    inp = synth.synth_inputs(B, K, seed=56 + r, device=dev)
    st = SilhouetteStep(B, K, n, dev)
    st.bind(inp["mus"].reshape(B, K, 3).contiguous().to(dev), inp["sigma"].contiguous().to(dev),
            *(inp[k].reshape(B).contiguous().to(dev) for k in ("rotx", "roty", "rotz")))
    gm = (-(torch.sign(torch.randn(B, n, n, device=dev))) / float(B * n * n)).contiguous()
    st.forward(); st.backward(gm)
    sets.append((st, gm))
torch.cuda.synchronize()
def step(s, gm, pf, sf, sb, pb):
    mus, sigma, rotx, roty, rotz = s._inputs
    stream = s._stream()
    if pf: check(lib.gg_project_fwd(mus.data_ptr(), sigma.data_ptr(), s.sigma_f64, rotx.data_ptr(), roty.data_ptr(), rotz.data_ptr(), *s.t, s.focal, s.B, s.K, None, None, s.params.data_ptr(), s.dev.index, stream), "pf")
    if sf: check(lib.gg_splat_fwd(s.params.data_ptr(), s.B, s.K, 1, s._sizes, s._null_ptrs, s._mask_ptrs, LAYOUT_NHWC, s.dev.index, stream), "sf")
    if sb: check(lib.gg_splat_bwd(s.params.data_ptr(), s.B, s.K, 1, s._sizes, s._null_ptrs, ptr_array([gm.data_ptr()]), LAYOUT_NHWC, s.partials.data_ptr(), s.T, s.dev.index, stream), "sb")
    if pb: check(lib.gg_project_bwd(mus.data_ptr(), sigma.data_ptr(), s.sigma_f64, rotx.data_ptr(), roty.data_ptr(), rotz.data_ptr(), *s.t, s.focal, s.B, s.K, s.partials.data_ptr(), s.T, None, None, s.dmus.data_ptr(), s.dsigma.data_ptr(), s.drot.data_ptr(), s.dev.index, stream), "pb")
def timeit(flags):
    def chain():
        for i in range(64):
            s, g = sets[i % R]; step(s, g, *flags)
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side): chain()
    torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): chain()
    for _ in range(10): gr.replay()
    torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50): gr.replay()
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 3200 * 1e3)
    return sorted(ts)[2]
for name, flags in [("full step", (1, 1, 1, 1)), ("no project_fwd", (0, 1, 1, 1)), ("no project_bwd", (1, 1, 1, 0)),
                    ("splats only", (0, 1, 1, 0)), ("splat_fwd only", (0, 1, 0, 0)), ("splat_bwd only", (0, 0, 1, 0)),
                    ("project_fwd only", (1, 0, 0, 0)), ("project_bwd only", (0, 0, 0, 1)), ("projections only", (1, 0, 0, 1)),
                    ("pf+sf", (1, 1, 0, 0)), ("sb+pb", (0, 0, 1, 1))]:
    print(f"{name:18s} {timeit(flags):7.2f} us per step")
