"""Launch list (ncu --metrics gpu__time_duration.sum --csv) -> per-kernel table: count, mean/min duration, share.
    python tools/summarize_launches.py gpurun_out/r2b_launches_bench.csv > profiles/r2b_launches_bench_summary.txt"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1]) as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
per = collections.OrderedDict()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"])[:90]
    per.setdefault(name, []).append(us)
tot = sum(sum(v) for v in per.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv python bench.py --steps 64 --warmup 3 --no-cpu")
print("(first 700 launches of the bench: the headline step chain; per-launch times are cold-cache and serialised -- compare SHARES)\n")
for k, v in per.items():
    print(f"{k:92s} n={len(v):4d}  mean {sum(v)/len(v):8.2f} us  min {min(v):8.2f}  share {100*sum(v)/tot:5.1f} %")
step = [k for k in per if any(s in k for s in ("project_fwd", "splat_fwd_mask_fast", "splat_bwd_mask_fast", "project_bwd"))]
s = sum(sum(per[k]) / len(per[k]) for k in step)
bw = [k for k in step if "splat_bwd" in k]
if bw:
    print(f"\nsum of the four kernels of one step (mean launch each): {s:.1f} us; splat_bwd share {100 * (sum(per[bw[0]])/len(per[bw[0]])) / s:.0f} %")
