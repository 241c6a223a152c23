"""Experiment: one C2 step (B=64) executed as `parts` independent sub-batch chains on parallel streams (fork/join inside
the step; consecutive steps stay serial).  Batch elements are independent, so this is an implementation
detail of the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussigan_b200.plan import SilhouetteStep
from gaussigan_b200 import synth
dev = torch.device("cuda:0")
B, K, n, R, ROUNDS = 64, 16, 256, 8, 4
def build(parts):
    sets = []
    for r in range(R):
        inp = synth.synth_inputs(B, K, seed=56 + r, device=dev)
        mus = inp["mus"].reshape(B, K, 3).contiguous().to(dev); sig = inp["sigma"].contiguous().to(dev)
        rx, ry, rz = (inp[k].reshape(B).contiguous().to(dev) for k in ("rotx", "roty", "rotz"))
        gm = (-(torch.sign(torch.randn(B, n, n, device=dev))) / float(B * n * n)).contiguous()
        sub = []
        for p in range(parts):
            lo, hi = p * B // parts, (p + 1) * B // parts
            st = SilhouetteStep(hi - lo, K, n, dev)
            st.bind(mus[lo:hi].contiguous(), sig[lo:hi].contiguous(), rx[lo:hi].contiguous(), ry[lo:hi].contiguous(), rz[lo:hi].contiguous())
            sub.append((st, gm[lo:hi].contiguous()))
        sets.append(sub)
    streams = [torch.cuda.Stream(dev) for _ in range(parts - 1)]
    def run_all():
        cur = torch.cuda.current_stream(dev)
        for _ in range(ROUNDS):
            for sub in sets:
                for s_ in streams: s_.wait_stream(cur)          # fork
                for p, (st, gm) in enumerate(sub):
                    with torch.cuda.stream(cur if p == 0 else streams[p - 1]):
                        st.forward(); st.backward(gm)
                for s_ in streams: cur.wait_stream(s_)          # join: the next step starts after the whole batch
    side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side): run_all()
    torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): run_all()
    return g, sets
for parts in (1, 2, 4):
    g, keep = build(parts)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(60): g.replay()
    b.record(); torch.cuda.synchronize()
    us = a.elapsed_time(b) / (60 * R * ROUNDS) * 1e3
    print(f"parts={parts}: {us:.2f} us per step of 64 silhouettes = {B / us * 1e6:,.0f} silhouettes/s")
    del g, keep
