"""Key metrics + top stall reasons of every kernel in an .ncu-rep as JSON (for profiles/).
   python scripts/ncu_to_json.py gpurun_out/X.ncu-rep profiles/X_summary.json "what was captured" """
import csv, io, json, subprocess, sys

rep, dst, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rr = list(csv.reader(io.StringIO(out)))
h, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct"]
kern = []
for r in rr[2:]:
    d = {"kernel": r[h.index("Kernel Name")].replace("<unnamed>::", "").split("(")[0]}
    for w in want:
        if w in h:
            d[w] = f"{r[h.index(w)]} {units[h.index(w)]}"
    st = sorted(((float(r[i].replace(',', '') or 0), hh) for i, hh in enumerate(h) if "issue_stalled" in hh and hh.endswith(".ratio")), reverse=True)[:6]
    d["top_stalls_per_issue"] = {hh.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""): round(v, 2) for v, hh in st}
    kern.append(d)
json.dump({"source": rep + " (ncu --set full --clock-control none)", "note": note, "kernels": kern}, open(dst, "w"), indent=1)
print(json.dumps(kern, indent=1)[:1500])
