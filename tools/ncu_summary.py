#!/usr/bin/env python3
"""Summary of one kernel launch of an `ncu --set full` report in the form kept under profiles/: the metrics DESIGN.md quotes.
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv ; python tools/ncu_summary.py raw.csv <elements> <algorithmic bytes per element> "<header line>" """
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
nEl = int(sys.argv[2]); algB = float(sys.argv[3]); head = sys.argv[4] if len(sys.argv) > 4 else ""
want = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
print(head); print()
d = {}
for i, h in enumerate(hdr):
    if h in want or h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        print("%-116s %-16s %s" % (h, units[i], vals[i]))
    d[h] = (units[i], vals[i])
def num(k):
    u, v = d[k]; v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
tr = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
print()
print("DRAM traffic per element: %.1f B (algorithmic %.0f B)" % (tr / nEl, algB))
print("shared-memory wavefronts per element: %.0f, of which bank-conflict replays %.1f %%" % (num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / nEl,
      100.0 * num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")))
print("warp instructions per element: %.0f" % (num("smsp__inst_executed.sum") / nEl))
