"""Turns gpurun_out/{launches_TAG.csv, prof_TAG.ncu-rep, bench_TAG.json} into the tracked summaries under profiles/."""
import collections, csv, io, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

# ---- launch list: per-kernel totals and share
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows:
    if r is hdr or len(r) <= iV or r[iK] == "Kernel Name":
        continue
    try:
        v = float(r[iV].replace(",", ""))
    except ValueError:
        continue
    k = r[iK].replace("<unnamed>::", "").split("(")[0]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, f"{tag}_launch_list_summary.csv"), "w") as f:
    f.write("kernel,launches,total_ns,avg_ns,share_of_captured_time\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"\"{k}\",{n},{t:.0f},{t / n:.0f},{t / tot:.4f}\n")
# the raw list, trimmed to kernel name + duration
with open(os.path.join(P, f"{tag}_launch_list.csv"), "w") as f:
    f.write("id,kernel,gpu__time_duration.sum_ns\n")
    for r in rows:
        if r is hdr or r[iK] == "Kernel Name":
            continue
        f.write(f"{r[0]},\"{r[iK].replace('<unnamed>::', '').split('(')[0]}\",{r[iV]}\n")

# ---- full capture: per-kernel metrics
out = subprocess.check_output(["ncu", "-i", os.path.join(G, f"prof_{tag}.ncu-rep"), "--page", "raw", "--csv"], text=True)
rr = list(csv.reader(io.StringIO(out)))
h, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
kern = []
for r in rr[2:]:
    d = {"kernel": r[h.index("Kernel Name")].replace("<unnamed>::", "").split("(")[0]}
    for w in want:
        if w in h:
            d[w] = f"{r[h.index(w)]} {units[h.index(w)]}"
    st = sorted(((float(r[i].replace(',', '') or 0), hh) for i, hh in enumerate(h) if "issue_stalled" in hh and hh.endswith(".ratio")), reverse=True)[:6]
    d["top_stalls_per_issue"] = {hh.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 2) for v, hh in st}
    kern.append(d)


def gb(s):
    v, u = s.split()
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]


cells = 128 * 700 * 700
dram = sum(gb(k["dram__bytes_read.sum"]) + gb(k["dram__bytes_write.sum"]) for k in kern)
summary = {"source": f"gpurun_out/prof_{tag}.ncu-rep (ncu --set full --clock-control none; 128 x 700^2 environments, one RK4 step; "
                     "ncu serialises the four k_fused_step variants of the launch set)",
           "cells_per_launch_set": cells, "dram_bytes_per_launch_set": dram, "dram_bytes_per_cell_update": dram / cells,
           "algorithmic_bytes_per_cell_update": 96, "kernels": kern}
json.dump(summary, open(os.path.join(P, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
json.dump({"source": summary["source"], "cells": cells, "dram_bytes_per_cell_update": dram / cells},
          open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
b = os.path.join(G, f"bench_{tag}.json")
if os.path.exists(b):
    open(os.path.join(P, f"{tag}_bench_line.json"), "w").write(open(b).read())
print(open(os.path.join(P, f"{tag}_launch_list_summary.csv")).read())
print(json.dumps({k: v for k, v in summary.items() if k != "kernels"}, indent=1))
