import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from gaussigan_b200.host import HostSilhouetteStep
from oracle import reference_port as rp
dev = torch.device('cuda:0'); B, K, n = 64, 16, 256
# raw H2D bandwidth
for mb in (4, 16, 64):
    h = torch.empty(mb * 2**20, dtype=torch.uint8).pin_memory(); d = torch.empty_like(h, device=dev)
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print(f"H2D {mb} MiB: {mb*2**20*20/(a.elapsed_time(b)*1e-3)/1e9:.1f} GB/s")
inp = rp.synth_inputs(B, K, seed=56); tgt = rp.synth_target_masks(B, n)
for depth in (1, 2, 3):
    host = HostSilhouetteStep(B, K, n, dev, depth=depth)
    hs = host.make_host_inputs(inp, tgt)
    for _ in range(5): host.submit(hs)
    host.drain(); torch.cuda.synchronize()
    t0 = time.perf_counter(); a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True); a.record()
    N = 300
    for _ in range(N): host.submit(hs)
    t1 = time.perf_counter(); host.drain(); b.record(); torch.cuda.synchronize()
    print(f"depth {depth}: {a.elapsed_time(b)/N*1e3:.1f} us/step device, host submit {(t1-t0)/N*1e6:.1f} us/step, loss {host.results[0].loss:.6f}")
# check loss/grads vs oracle
mask_ref, loss_ref, dmu_ref, dsig_ref, drot_ref = rp.render_loss_and_grads(inp, tgt, n)
res = host.submit(hs).wait()
print('loss', res.loss, float(loss_ref), 'dmus rel', float((res.dmus - dmu_ref.reshape(B,K,3)).abs().max()/dmu_ref.abs().max()), 'dsig rel', float((res.dsigma - dsig_ref).abs().max()/dsig_ref.abs().max()))
