#!/usr/bin/env python3
"""Selected metrics of every kernel in an `ncu -i X.ncu-rep --page raw --csv` export, one column per launch:
   ncu -i gpurun_out/X.ncu-rep --page raw --csv > /tmp/x.csv ; python tools/ncu_summary.py /tmp/x.csv label1,label2,... > profiles/..csv"""
import csv, sys
METRICS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
labels = sys.argv[2].split(",") if len(sys.argv) > 2 else [r[hdr.index("Kernel Name")][:40] for r in data]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit"] + labels)
for m in METRICS:
    if m in hdr:
        i = hdr.index(m)
        w.writerow([m, units[i]] + [r[i] for r in data])
