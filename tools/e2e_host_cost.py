"""Host-side cost of one HostSilhouetteStep.submit and the effect of the slot depth (development tool): wall clock per
submit with the real batch and with a batch of one silhouette (GPU work negligible -> what remains is Python + ctypes +
the CUDA API calls)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.host import HostSilhouetteStep
from gaussigan_b200 import synth
dev = torch.device("cuda:0")
depths = [int(x) for x in os.environ.get("DEPTHS", "6").split(",")]
for B in (64, 1):
    for depth in depths:
        K, n = 16, 256
        host = HostSilhouetteStep(B, K, n, dev, depth=depth, target_strips=True)
        inp = synth.synth_inputs(B, K, seed=56, device=dev)
        tgt = synth.synth_target_masks(B, n)
        h = host.make_host_inputs(inp, tgt)
        for _ in range(1500): host.submit(h)
        host.drain(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        N = 4000
        for _ in range(N): host.submit(h)
        t1 = time.perf_counter()
        host.drain(); torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"B={B} depth={depth}: submit loop {1e6*(t1-t0)/N:.1f} us per submit (host side), {1e6*(t2-t0)/N:.1f} us per step incl. drain")
