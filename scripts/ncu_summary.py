"""Summarises an .ncu-rep (from `ncu --set full`) into a small markdown table for profiles/."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs"), ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "lanes/inst"),
]
idx = [(hdr.index(k), label) for k, label in want if k in hdr]
with open(out, "w") as f:
    f.write(f"# ncu --set full summary of `{rep}`\n\n")
    f.write("| " + " | ".join(label for _, label in idx) + " |\n")
    f.write("|" + "---|" * len(idx) + "\n")
    for r in rows[2:]:
        cells = []
        for i, label in idx:
            v = r[i]
            if label == "kernel":
                v = v.split("(")[0]
            elif units[i] and label in ("time", "dram read", "dram write"):
                v = f"{v} {units[i]}"
            cells.append(v)
        f.write("| " + " | ".join(cells) + " |\n")
print(open(out).read())
