"""The bench's headline step (8 rotating buffer sets, 64-step CUDA-graph chain) timed alone: quick A/B of launch knobs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.plan import SilhouetteStep, SilhouetteLossStep
from gaussigan_b200 import synth
dev = torch.device("cuda:0"); B, K, n, R = 64, 16, 256, 8
sets = []
for r in range(R):
    inp = synth.synth_inputs(B, K, seed=56 + r, device=dev)
    st = SilhouetteStep(B, K, n, dev)
    st.bind(inp["mus"].reshape(B, K, 3).contiguous().to(dev), inp["sigma"].contiguous().to(dev),
            *(inp[k].reshape(B).contiguous().to(dev) for k in ("rotx", "roty", "rotz")))
    gm = (-(torch.sign(torch.randn(B, n, n, device=dev))) / float(B * n * n)).contiguous()
    sets.append((st, gm))
def chain():
    for i in range(64):
        s, g = sets[i % R]; s.forward(); s.backward(g)
side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side): chain()
torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr): chain()
for _ in range(20): gr.replay()
torch.cuda.synchronize()
ts = []
for rep in range(5):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(100): gr.replay()
    b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) / 6400 * 1e3)
print(f"{os.environ.get('TAG','')}: step {min(ts):.2f} us (best of 5 x 6400 steps; median {sorted(ts)[2]:.2f})")
