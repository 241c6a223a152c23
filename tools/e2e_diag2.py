"""e2e step rate vs pipeline depth, in one process (boxes differ a lot in PCIe behaviour: compare within a run)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.host import HostSilhouetteStep
from gaussigan_b200 import synth as rp
dev = torch.device('cuda:0'); B, K, n = 64, 16, 256
inp = rp.synth_inputs(B, K, seed=56); tgt = rp.synth_target_masks(B, n)
h = torch.empty(4281088, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device=dev)
for _ in range(3): d.copy_(h, non_blocking=True)
torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): d.copy_(h, non_blocking=True)
b.record(); torch.cuda.synchronize()
print(f"H2D 4.28 MB alone: {a.elapsed_time(b)/50*1e3:.1f} us")
for depth in (2, 3, 4, 6, 8):
    host = HostSilhouetteStep(B, K, n, dev, depth=depth)
    hs = [host.make_host_inputs(inp, tgt) for _ in range(2)]
    for i in range(10): host.submit(hs[i % 2])
    host.drain(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    N = 400
    for i in range(N): host.submit(hs[i % 2])
    host.drain(); b.record(); torch.cuda.synchronize()
    print(f"depth {depth}: {a.elapsed_time(b)/N*1e3:.1f} us/step  loss {host.results[0].loss:.6f}")
