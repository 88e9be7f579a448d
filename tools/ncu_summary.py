"""Summarise an .ncu-rep (ncu --set full) per launch: duration, DRAM bytes, fp64 pipe, occupancy, top stall reasons.
python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

rows = list(csv.reader(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
idx = [hdr.index(w) for w in want]
stall = [i for i, h in enumerate(hdr) if "issue_stalled" in h and "per_issue_active" in h and "not_issued" not in h]
print("units:", [rows[1][i] for i in idx])
for r in rows[2:]:
    print(r[hdr.index("Kernel Name")][:60])
    print("   ", " | ".join(f"{w.split('.')[0].replace('__', ':')[-28:]}={r[i][:10]}" for w, i in zip(want, idx)))
    v = sorted(((float(r[i]), hdr[i].split("issue_stalled_")[1].split("_per_issue")[0]) for i in stall), reverse=True)[:6]
    print("    stalls:", ", ".join(f"{n}={x:.2f}" for x, n in v))
