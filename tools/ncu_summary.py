#!/usr/bin/env python
"""One line per profiled kernel from an .ncu-rep (ncu --page raw --csv): time, regs, occupancy, FP64 pipe, DRAM."""
import csv, subprocess, sys, re
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, data = rows[0], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=float("nan")):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return d
seen = {}
print(f"{'kernel':58s} {'us':>8s} {'regs':>4s} {'warps%':>6s} {'fp64%':>6s} {'issue%':>6s} {'l1%':>5s} {'dramGB/s':>9s} {'dramMB':>8s} {'inst/1e6':>8s}")
for r in data:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "")
    grid = r[ix.get("Grid Size", 0)] if "Grid Size" in ix else ""
    key = (name, grid)
    if key in seen and "--all" not in sys.argv: continue
    seen[key] = 1
    us = g(r, "gpu__time_duration.sum")
    unit = rows[1][ix["gpu__time_duration.sum"]]
    if unit == "ns": us /= 1e3
    elif unit == "ms": us *= 1e3
    def bytes_of(k):
        v = g(r, k); u = rows[1][ix[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    db = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
    print(f"{name[:58]:58s} {us:8.1f} {g(r,'launch__registers_per_thread'):4.0f} {g(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(r,'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):6.1f} {g(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{g(r,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} {db/us/1e3:9.1f} {db/1e6:8.1f} {g(r,'smsp__inst_executed.sum')/1e6:8.2f}")
    if "--stalls" in sys.argv:
        st = {h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): g(r, h) for h in hdr
              if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")}
        top = sorted(st.items(), key=lambda kv: -kv[1])[:7]
        print("      stalls/issue: " + "  ".join(f"{k}={v:.2f}" for k, v in top) +
              f"   | eligible/cycle={g(r,'smsp__warps_eligible.avg.per_cycle_active'):.2f} active warps/smsp={g(r,'smsp__warps_active.avg.per_cycle_active'):.2f}"
              f" local ld/st={g(r,'smsp__inst_executed_op_local_ld.sum',0)+g(r,'smsp__inst_executed_op_local_st.sum',0):.0f}")
