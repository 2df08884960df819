"""Reads `ncu --page raw --csv` on stdin and prints the metrics the roofline section needs, per profiled launch."""
import csv
import sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_utcmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]
rd = csv.reader(sys.stdin)
hdr = next(rd)
units = next(rd)
col = {h: i for i, h in enumerate(hdr)}
tens = [h for h in hdr if "tensor" in h and "pct" in h]
for row in rd:
    if len(row) < len(hdr):
        continue
    print("----")
    for k in KEYS + [t for t in tens if t not in KEYS]:
        if k in col:
            print(f"{k} = {row[col[k]]} {units[col[k]]}")
