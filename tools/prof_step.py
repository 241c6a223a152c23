"""Tiny driver for ncu: a few eager (non-graph) forward+backward steps of the C2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.plan import SilhouetteStep
from gaussigan_b200 import synth as rp
B, K, n = int(os.environ.get("PB", 64)), 16, 256
dev = torch.device("cuda:0")
inp = rp.synth_inputs(B, K, seed=56)
step = SilhouetteStep(B, K, n, dev)
step.bind(inp["mus"].reshape(B, K, 3).contiguous().to(dev), inp["sigma"].to(dev), *(inp[k].reshape(B).contiguous().to(dev) for k in ("rotx", "roty", "rotz")))
gm = (-(torch.sign(torch.randn(B, n, n, device=dev))) / float(B * n * n)).contiguous()
for _ in range(int(os.environ.get("PSTEPS", 6))):
    step.forward(); step.backward(gm)
torch.cuda.synchronize()
print("ok")
if os.environ.get("PFUSED"):
    from gaussigan_b200.plan import SilhouetteLossStep
    fs = SilhouetteLossStep(B, K, n, dev)
    fs.bind(*step._inputs)
    tgt = rp.synth_target_masks(B, n)
    raw = ((tgt.reshape(B, n, n) + 1.0) * 127.5).round().clamp(0, 255).to(torch.uint8).to(dev)
    for _ in range(int(os.environ.get("PSTEPS", 6))):
        fs.loss_backward(raw)
    torch.cuda.synchronize()
    print("fused ok", fs.loss())
