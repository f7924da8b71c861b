"""Summarise `ncu -i rep --page raw --csv` into a markdown table of the metrics the roofline uses.
usage: ncu -i prof.ncu-rep --page raw --csv > raw.csv; python scripts/summarise_ncu.py raw.csv labels..."""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
rows = list(csv.reader(open(sys.argv[1])))
labels = sys.argv[2:]
hdr, units = rows[0], rows[1]
ki = hdr.index("Kernel Name")
print("| metric | " + " | ".join(labels[i] if i < len(labels) else r[ki][:40] for i, r in enumerate(rows[2:])) + " |")
print("|---|" + "---:|" * len(rows[2:]))
for m, nice in WANT:
    if m not in hdr:
        continue
    i = hdr.index(m)
    vals = []
    for r in rows[2:]:
        try:
            vals.append(f"{float(r[i].replace(',', '')):.4g} {units[i]}")
        except ValueError:
            vals.append(r[i])
    print(f"| {nice} (`{m}`) | " + " | ".join(vals) + " |")
