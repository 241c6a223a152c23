"""H2D of one step's target masks (4.19 MB): one copy vs 2/4 concurrent pieces on separate streams (noise: repeat)."""
import torch, time
dev = torch.device('cuda:0')
N = 4194304
h = torch.empty(N, dtype=torch.uint8).pin_memory(); d = torch.empty(N, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream(dev) for _ in range(4)]
def run(parts, reps=100):
    sz = N // parts
    torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    a.record()
    for r in range(reps):
        for p in range(parts):
            st = streams[p]
            st.wait_stream(cur) if r == 0 else None
            with torch.cuda.stream(st):
                d[p*sz:(p+1)*sz].copy_(h[p*sz:(p+1)*sz], non_blocking=True)
    for p in range(parts): cur.wait_stream(streams[p])
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for rnd in range(3):
    print("round", rnd, {p: round(run(p), 1) for p in (1, 2, 4)}, "us per 4 MiB")
