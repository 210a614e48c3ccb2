#!/usr/bin/env python
"""Summarise one `ncu --set full` report: `ncu -i X.ncu-rep --page raw --csv | python tools/ncu_summary.py [out.csv]`."""
import csv
import sys

WANT = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warp_issue_stalled", "sm__inst_executed_pipe_xu", "sm__pipe_xu", "sm__inst_executed_pipe_uniform")
rows = list(csv.reader(sys.stdin))
hdr, units, vals = rows[0], rows[1], rows[2:]
out = []
for r in vals:
    name = r[hdr.index("Kernel Name")]
    out.append(("kernel", "", name))
    for i, h in enumerate(hdr):
        if h.startswith(WANT) or (("tensor" in h) and h.endswith((".avg.pct_of_peak_sustained_elapsed", ".avg.pct_of_peak_sustained_active"))
                                  and "subpipe" not in h):
            out.append((h, units[i], r[i]))
w = csv.writer(open(sys.argv[1], "w", newline="") if len(sys.argv) > 1 else sys.stdout)
w.writerow(("metric", "unit", "value"))
w.writerows(out)
