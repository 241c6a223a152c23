"""Tiny driver for ncu: a few eager (non-graph) forward+backward steps of the C2 shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.plan import SilhouetteStep
from oracle import reference_port as rp
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
