#!/usr/bin/env python3
"""Key counters of `ncu --set full` captures (the `--page raw --csv` export): one block per captured launch."""
import csv
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    if len(rows) < 3:
        print("== %s: empty" % path)
        continue
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for d in rows[2:]:
        name = d[col["Kernel Name"]]
        print("== %s\n   %s" % (path.split("/")[-1], name[:150]))
        rd = wr = dur = None
        for k in KEYS:
            if k in col:
                v, u = d[col[k]], units[col[k]]
                print("   %-88s %14s %s" % (k, v, u))
                if k == "dram__bytes_read.sum":
                    rd = float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
                if k == "dram__bytes_write.sum":
                    wr = float(v) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}.get(u, 1)
                if k == "gpu__time_duration.sum":
                    dur = float(v) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(u, 1e-6)
        if rd is not None and wr is not None and dur:
            print("   %-88s %14.1f GB/s (traffic %.1f MB per launch)" % ("dram read+write / duration", (rd + wr) / dur / 1e9, (rd + wr) / 1e6))
