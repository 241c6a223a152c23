"""Run under torchrun on N GPUs: the NVLink peer-memory reduction of the shared gradients against NCCL, values
and latency.  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_allreduce_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gaussigan_b200.dist import P2PSharedGradReducer, reduce_shared_grads
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, K = 64, 16
g = torch.Generator().manual_seed(100 + rank)
red = P2PSharedGradReducer(K, dev)
ok = True
for it in range(20):
    dmus = torch.randn(B, K, 3, generator=g).to(dev); dsig = torch.randn(B, K, 3, 3, generator=g, dtype=torch.float64).to(dev)
    a, b = red(dmus, dsig)
    a, b = a.clone(), b.clone()
    ra, rb = reduce_shared_grads([dmus.double(), dsig])
    red.check()
    ea = float((a - ra.view(K, 3)).abs().max() / ra.abs().max()); eb = float((b - rb.view(K, 3, 3)).abs().max() / rb.abs().max())
    ok = ok and ea < 1e-13 and eb < 1e-13
    # identical bits on every rank
    ref = a.clone(); dist.broadcast(ref, 0)
    ok = ok and bool(torch.equal(ref, a))
dmus = torch.randn(B, K, 3, generator=g).to(dev); dsig = torch.randn(B, K, 3, 3, generator=g, dtype=torch.float64).to(dev)
def timeit(fn, n=200):
    for _ in range(20): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
t_p2p = timeit(lambda: red(dmus, dsig)); red.check()
t_nccl = timeit(lambda: reduce_shared_grads([dmus, dsig]))
# graph replay of the P2P kernel
gr = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side): red(dmus, dsig)
torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize(); dist.barrier()
with torch.cuda.graph(gr): red(dmus, dsig)
t_graph = timeit(lambda: gr.replay()); red.check()
res = torch.tensor([t_p2p, t_nccl, t_graph, float(ok)], device=dev, dtype=torch.float64)
dist.all_reduce(res, op=dist.ReduceOp.MAX if True else None)
okall = torch.tensor([float(ok)], device=dev); dist.all_reduce(okall, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}: values match NCCL and are bit-identical across ranks: {bool(okall.item())}; "
          f"per call (max over ranks): P2P kernel {res[0]:.1f} us, P2P kernel in a CUDA graph {res[2]:.1f} us, local sum + NCCL all_reduce {res[1]:.1f} us")
dist.destroy_process_group()
