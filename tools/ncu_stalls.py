"""Per-instruction warp-stall samples of one kernel from an ncu report captured with --import-source on.
usage: ncu -i rep.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv; python tools/ncu_stalls.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Address")
data = []
for r in rows[rows.index(hdr) + 1:]:          # first kernel section only
    if r and r[0] in ("Kernel Name", "Address"):
        break
    if len(r) == len(hdr):
        data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
print({k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
    st = {s[6:]: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:4])
    print(r[ix["Address"]][-4:], r[ix["Source"]].strip()[:64].ljust(64), r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(8), st)
