"""Condense an .ncu-rep into the handful of numbers we track (run where ncu is installed):
   python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/ncu/name.txt"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEEP = [r"^gpu__time_duration\.sum$", r"^sm__cycles_elapsed\.avg$", r"^sm__cycles_elapsed\.avg\.per_second$",
        r"^sm__cycles_active\.avg$", r"^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|cluster_x|occupancy_limit_\w+)$",
        r"^sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$",
        r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
        r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
        r"^sm__inst_executed_pipe_xu\.avg\.pct_of_peak_sustained_active$",
        r"^sm__pipe_(fma|alu)_cycles_active\.avg\.pct_of_peak_sustained_active$",
        r"^dram__bytes_(read|write)\.sum$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
        r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__t_sector_hit_rate\.pct$",
        r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$",
        r"^l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum$"]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))
    print("kernel:", d.get("Kernel Name", "?")[:100])
    for h in hdr:
        if any(re.search(k, h) for k in KEEP):
            print(f"  {h:75s} {d[h]} {u[h]}")
    stalls = []
    for h in hdr:
        m = re.match(r"smsp__pcsamp_warps_issue_stalled_(\w+)$", h)
        if m and not m.group(1).endswith("not_issued"):
            try:
                stalls.append((float(d[h]), m.group(1)))
            except ValueError:
                pass
    tot = sum(x for x, _ in stalls) or 1.0
    print("  warp stall samples (pc sampling):")
    for x, n in sorted(stalls, reverse=True)[:8]:
        print(f"    {n:28s} {100 * x / tot:5.1f} %")
