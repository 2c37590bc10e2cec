"""Developer aid: print the key metrics of every kernel in an .ncu-rep (ncu --page raw --csv)."""
import csv, io, subprocess, sys
out = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__cycles_elapsed.avg', 'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active']
extra = [h for h in hdr if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct')]
for r in rows[2:]:
    print('-' * 80)
    for w in want:
        if w in hdr:
            print(f"{w:70s} {r[hdr.index(w)]} {units[hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(h)] or 0), h) for h in extra), reverse=True)[:8]
    for v, h in st:
        print(f"   stall {h.replace('smsp__average_warp','').replace('_per_warp_active.pct',''):60s} {v:.1f}")
