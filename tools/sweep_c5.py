"""BASELINE configs[4]: throughput sweep, batch 4096, K=64, 512x512, mask fwd+bwd, batch-sharded over the ranks
(strong scaling: the total batch is fixed).  Launch: python tools/sweep_c5.py            (1 GPU)
        python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep_c5.py   (N GPUs)
One JSON line from rank 0; device time with CUDA events, barrier on both sides, max over ranks."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gaussigan_b200.plan import SilhouetteStep
from gaussigan_b200.dist import shard_range
from gaussigan_b200 import synth as rp

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
BT, K, n = int(os.environ.get("C5_BATCH", 4096)), 64, 512
lo, hi = shard_range(BT, world, rank); B = hi - lo
base = rp.synth_inputs(256, K, seed=56 + rank)
rep = (B + 255) // 256
f = lambda t: t.repeat(rep, *([1] * (t.dim() - 1)))[:B].contiguous().to(dev)
step = SilhouetteStep(B, K, n, dev)
step.bind(f(base["mus"].reshape(-1, K, 3)), f(base["sigma"]), f(base["rotx"].reshape(-1)), f(base["roty"].reshape(-1)), f(base["rotz"].reshape(-1)))
gm = torch.randn(B, n, n, device=dev) / (BT * n * n)
step.capture(gm)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize(dev)
for _ in range(3): step.graph.replay()
barrier()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
steps = int(os.environ.get("C5_STEPS", 10))
a.record()
for _ in range(steps): step.graph.replay()
b.record(); barrier()
ms = a.elapsed_time(b) / steps
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
if rank == 0:
    pxk = BT * n * n * K
    print(json.dumps({"workload": f"C5: mask fwd+bwd, batch {BT}, K={K}, {n}x{n}, batch-sharded x{world} ({B} per GPU), no data-path collective",
                      "n_gpus": world, "ms_per_step": ms, "silhouettes_per_s": BT / (ms * 1e-3), "scaling": "strong",
                      "T_pixel_gaussians_per_s": pxk / (ms * 1e-3) / 1e12,
                      "frac_of_26_slot_issue_roofline": 26 * pxk / (ms * 1e-3) / (world * 148 * 128 * 1.965e9)}))
if world > 1: dist.destroy_process_group()
