"""Condense an .ncu-rep (ncu --set full) into the few numbers DESIGN.md / profiles/ quote.
Usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-name-substring] > profiles/x_ncu_summary.txt"""
import csv
import io
import subprocess
import sys

KEEP = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__cluster_size",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__warps_active.avg.per_cycle_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg"]

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
print(f"# ncu --set full --clock-control none --import-source on ; source: {rep} (B200)")
for r in rows[2:]:
    rec = dict(zip(hdr, r))
    if flt and flt not in rec.get("Kernel Name", ""):
        continue
    for i, h in enumerate(hdr):
        if any(h == k or h.endswith(k) for k in KEEP):
            print(f"{h} [{units[i]}] = {r[i]}")
    print("--")
