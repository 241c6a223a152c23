"""Run under torchrun on N GPUs: the NVLink peer-memory reduction of the shared gradients against NCCL -- values,
bit-identity across ranks, a delayed rank, the post/wait split, a timed-out peer, latency.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/p2p_allreduce_check.py [--timeout-case]
Prints one line per check from rank 0 and exits non-zero if any check fails (tests/test_dist_gpu.py runs it)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gaussigan_b200.dist import P2PSharedGradReducer, reduce_shared_grads
world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, K = 64, 16
g = torch.Generator().manual_seed(100 + rank)
red = P2PSharedGradReducer(K, dev)
fails = []


def all_ok(flag):
    t = torch.tensor([float(flag)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN); return bool(t.item())


def rand_grads():
    return torch.randn(B, K, 3, generator=g).to(dev), torch.randn(B, K, 3, 3, generator=g, dtype=torch.float64).to(dev)


def matches_nccl(a, b, dmus, dsig):
    ra, rb = reduce_shared_grads([dmus.double(), dsig])
    ea = float((a - ra.view(K, 3)).abs().max() / ra.abs().max()); eb = float((b - rb.view(K, 3, 3)).abs().max() / rb.abs().max())
    ref = a.clone(); dist.broadcast(ref, 0)                       # identical bits on every rank
    return ea < 1e-13 and eb < 1e-13 and bool(torch.equal(ref, a))


if "--timeout-case" in sys.argv:
    # GG_P2P_TIMEOUT_MS is small (set by the caller): rank 1 never posts call 2 in time -> every other rank's result is NaN
    # and check() raises; the late rank itself gets a correct sum (its peers did post).
    dmus, dsig = rand_grads()
    a, b = red(dmus, dsig); red.check()                          # call 1: everybody on time
    ok1 = matches_nccl(a.clone(), b.clone(), dmus, dsig)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 1:
        time.sleep(3.0)
    a, b = red(dmus, dsig)
    torch.cuda.synchronize()
    raised = False
    try:
        red.check()
    except RuntimeError:
        raised = True
    nan = bool(torch.isnan(a).all() and torch.isnan(b).all())
    good = (not raised and not nan) if rank == 1 else (raised and nan)
    ok = all_ok(ok1 and good)
    if rank == 0:
        print(f"world {world}: timed-out peer -> NaN results and a latched error on the waiting ranks: {ok}")
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if ok else 1)

ok = True
for it in range(20):
    dmus, dsig = rand_grads()
    a, b = red(dmus, dsig)
    a, b = a.clone(), b.clone()
    red.check()
    ok = ok and matches_nccl(a, b, dmus, dsig)
ok = all_ok(ok)
if rank == 0: print(f"world {world}: 20 calls match NCCL and are bit-identical across ranks: {ok}")
fails.append(not ok)

# a rank that arrives 1.5 s late (checkpoint save, data-loader stall ...): the others wait, nobody gets stale sums
dmus, dsig = rand_grads()
torch.cuda.synchronize(); dist.barrier()
if rank == world - 1:
    time.sleep(1.5)
a, b = red(dmus, dsig); a, b = a.clone(), b.clone(); red.check()
ok = all_ok(matches_nccl(a, b, dmus, dsig))
if rank == 0: print(f"world {world}: a rank delayed by 1.5 s still yields the correct sum everywhere: {ok}")
fails.append(not ok)

# post ... other work ... wait
ok = True
for it in range(5):
    dmus, dsig = rand_grads()
    red.post(dmus, dsig)
    filler = torch.randn(1 << 22, device=dev).sin().sum()
    a, b = red.wait(); a, b = a.clone(), b.clone(); red.check()
    ok = ok and matches_nccl(a, b, dmus, dsig) and bool(torch.isfinite(filler))
ok = all_ok(ok)
if rank == 0: print(f"world {world}: post / wait with work in between: {ok}")
fails.append(not ok)

dmus, dsig = rand_grads()
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
# graph replay of post + wait
gr = torch.cuda.CUDAGraph()
side = torch.cuda.Stream(dev); side.wait_stream(torch.cuda.current_stream(dev))
with torch.cuda.stream(side): red(dmus, dsig)
torch.cuda.current_stream(dev).wait_stream(side); torch.cuda.synchronize(); dist.barrier()
with torch.cuda.graph(gr): red(dmus, dsig)
t_graph = timeit(lambda: gr.replay()); red.check()
res = torch.tensor([t_p2p, t_nccl, t_graph], device=dev, dtype=torch.float64)
dist.all_reduce(res, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world {world}: per call (max over ranks): post+wait eager {res[0]:.1f} us, post+wait in a CUDA graph {res[2]:.1f} us, "
          f"local sum + NCCL all_reduce {res[1]:.1f} us")
dist.barrier(); dist.destroy_process_group()
sys.exit(1 if any(fails) else 0)
