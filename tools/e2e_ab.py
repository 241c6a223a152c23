"""e2e step rate of the host-buffer step in this process' configuration (env), several repetitions."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.host import HostSilhouetteStep
from gaussigan_b200 import synth as rp
dev = torch.device('cuda:0'); B, K, n = 64, 16, 256
depth = int(os.environ.get("DEPTH", 4))
inp = rp.synth_inputs(B, K, seed=56); tgt = rp.synth_target_masks(B, n)
host = HostSilhouetteStep(B, K, n, dev, depth=depth)
hs = [host.make_host_inputs(inp, tgt) for _ in range(4)]
for i in range(300): host.submit(hs[i % 4])
host.drain(); torch.cuda.synchronize()
out = []
for rep in range(4):
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    N = 500
    for i in range(N): host.submit(hs[i % 4])
    host.drain(); b.record(); torch.cuda.synchronize()
    out.append(round(a.elapsed_time(b) / N * 1e3, 1))
print(f"ZEROCOPY={os.environ.get('GG_ZEROCOPY','1')} ZC_TARGET={os.environ.get('GG_ZEROCOPY_TARGET','0')} depth={depth}: us/step {out}")
