"""Config 3 (BASELINE configs[2]): dynamic-object step, B=256, K=16, 256^2 -- per-frame deformation + render + fused
L1 loss + full backward to the deformation heads and the canonical frame, through the public autograd API."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gaussigan_b200 as gg
from gaussigan_b200 import synth as rp
dev = torch.device("cuda:0")
B, K, n = int(os.environ.get("B", 256)), 16, 256
g = torch.Generator().manual_seed(33)
th = lambda *s: torch.tanh(torch.randn(*s, generator=g))
frame = [th(1, K, 3).to(dev).requires_grad_(True), th(1, K, 3).to(dev).requires_grad_(True), torch.randn(1, K, 3, generator=g).to(dev).requires_grad_(True)]
cmu = th(1, K, 1, 3).to(dev).requires_grad_(True)
heads = [t.to(dev).requires_grad_(True) for t in (th(B, K * 3), th(B, K * 3), th(B, K), th(B, K), th(B, K), th(B, 1))]
tgt = rp.synth_target_masks(B, n, seed=4).to(dev)
def step():
    csig = gg.sigma_from_frame(*frame)
    mus, sig, theta = gg.deform(*heads, cmu, csig)
    z = torch.zeros_like(theta)
    loss = gg.mask_recon_loss_bis(tgt, mus, sig, z, theta, z, 0, 0, -2., size=n)
    loss.backward()
    return loss
for _ in range(5): step()
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
N = 50
for _ in range(N): step()
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / N * 1e3
print(f"C3 B={B}: {us:.1f} us/step = {B / us * 1e6:.0f} silhouettes/s (autograd API, eager)")
# the same step captured into one CUDA graph (gradients accumulate into static .grad tensors)
try:
    for t in frame + [cmu] + heads: t.grad = torch.zeros_like(t)
    def gstep():
        for t in frame + [cmu] + heads: t.grad.zero_()
        return step()
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        for _ in range(3): gstep()
    torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        loss_static = gstep()
    for _ in range(5): graph.replay()
    torch.cuda.synchronize()
    a.record()
    for _ in range(N): graph.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / N * 1e3
    print(f"C3 B={B}: {us:.1f} us/step = {B / us * 1e6:.0f} silhouettes/s (whole step in one CUDA graph), loss {float(loss_static):.6f}")
except Exception as exc:
    print("graph capture of the C3 step failed:", repr(exc))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))

# ---- the same step as a planned launch sequence (gaussigan_b200.plan.DynamicObjectStep), one CUDA graph
from gaussigan_b200.plan import DynamicObjectStep
gth = torch.Generator().manual_seed(5)
tt = lambda *s: torch.tanh(torch.randn(*s, generator=gth)).to(dev)
ps = DynamicObjectStep(B, K, n, dev)
ps.bind(tt(K, 3), tt(K, 3), torch.randn(K, 3, generator=gth).to(dev), tt(K, 3), tt(B, K, 3), tt(B, K, 3), tt(B, K, 3), tt(B))
raw = ((tgt.reshape(B, n, n) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8)
ps.capture(raw)
for _ in range(5): ps.graph.replay()
torch.cuda.synchronize()
a.record()
for _ in range(N): ps.graph.replay()
b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / N * 1e3
print(f"C3 B={B}: {us:.1f} us/step = {B / us * 1e6:.0f} silhouettes/s (DynamicObjectStep, planned launches in one CUDA graph), loss {ps.loss():.6f}")
