"""Turn .ncu-rep files into the committed text summaries under profiles/ (runs on the CPU box: `ncu -i`)."""
import csv, subprocess, sys, collections, json

def details(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); h = rows[0]; idx = {n: i for i, n in enumerate(h)}
    byid = collections.OrderedDict()
    for r in rows[1:]:
        d = byid.setdefault(r[idx["ID"]], {"name": r[idx["Kernel Name"]]})
        d[r[idx["Metric Name"]]] = (r[idx["Metric Value"]], r[idx["Metric Unit"]])
    return byid

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines())); h = rows[0]; units = dict(zip(h, rows[1]))
    return [dict(zip(h, r)) for r in rows[2:]], units

WANT = ["Duration", "SM Frequency", "Elapsed Cycles", "SM Active Cycles", "Registers Per Thread", "Grid Size", "Block Size",
        "Waves Per SM", "Theoretical Occupancy", "Achieved Occupancy", "Executed Ipc Active", "Issue Slots Busy",
        "Executed Instructions", "Warp Cycles Per Issued Instruction", "Eligible Warps Per Scheduler",
        "Active Warps Per Scheduler", "DRAM Throughput", "Memory Throughput", "L2 Hit Rate",
        "Static Shared Memory Per Block", "Dynamic Shared Memory Per Block"]
RAW = ["dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]
STALL = "smsp__average_warps_issue_stalled_"

def main(reps, out_txt, out_json=None):
    traffic = {}
    with open(out_txt, "w") as fh:
        for rep in reps:
            det = details(rep); rw, units = raw(rep)
            fh.write(f"##### {rep}  (ncu --set full --clock-control none --import-source on; one launch each)\n")
            for (kid, d), r in zip(det.items(), rw):
                fh.write(f"\n=== {d['name']}\n")
                for k in WANT:
                    if k in d: fh.write(f"  {k:42s} {d[k][0]:>14s} {d[k][1]}\n")
                for k in RAW:
                    if k in r: fh.write(f"  {k:72s} {r[k]} {units.get(k, '')}\n")
                st = {k[len(STALL):].replace("_per_issue_active.ratio", ""): float(v) for k, v in r.items()
                      if k.startswith(STALL) and k.endswith("per_issue_active.ratio") and v not in ("", "n/a")}
                top = sorted(st.items(), key=lambda kv: -kv[1])[:7]
                fh.write("  warp-state cycles per issued instruction: " + ", ".join(f"{k}={v:.2f}" for k, v in top) + "\n")
                try:
                    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                    traffic[d["name"].split("(")[0]] = int(sum(float(r[k].replace(",", "")) * mult[units[k]]
                                                               for k in ("dram__bytes_read.sum", "dram__bytes_write.sum")))
                except Exception:
                    pass
    if out_json:
        json.dump(traffic, open(out_json, "w"), indent=1)
    print(open(out_txt).read()[:3000])

if __name__ == "__main__":
    main(sys.argv[3:], sys.argv[1], sys.argv[2] if sys.argv[2] != "-" else None)
