"""Opcode histogram and hottest SASS lines of one kernel from `ncu -i REP --page source --csv` (development tool)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[start]; idx = {n: i for i, n in enumerate(h)}
data = []
for r in rows[start + 1:]:
    if len(r) <= idx["Instructions Executed"] or not r[0].startswith("0x"): break
    data.append((int(r[idx["Instructions Executed"]] or 0), int(r[idx["# Samples"]] or 0), r[idx["Source"]].strip()))
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
print(f"{len(data)} SASS lines, {tot} warp instructions executed, {ts} samples")
op = collections.Counter(); sm = collections.Counter()
for n, s, src in data:
    parts = src.split()
    o = parts[1] if parts and parts[0].startswith("@") else (parts[0] if parts else "")
    o = o.split(".")[0]
    op[o] += n; sm[o] += s
for o, n in op.most_common(top):
    print(f"  {o:14s} {n:11d} {n / tot * 100:5.1f}%   samples {sm[o] / max(ts, 1) * 100:5.1f}%")
print("hottest lines by samples:")
for n, s, src in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"  {s:6d} {n:9d}  {src[:110]}")
