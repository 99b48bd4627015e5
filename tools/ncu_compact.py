#!/usr/bin/env python
"""Dev tool: compact `ncu -i x.ncu-rep --page raw --csv` (one kernel, wide format) into the metric,unit,value list kept
under profiles/ (the counters the DESIGN/VERDICT discussion uses).    python tools/ncu_compact.py raw.csv > out.csv"""
import csv
import re
import sys

KEEP = re.compile(r"^(Kernel Name|dram__bytes_(read|write)\.sum|gpu__time_duration\.sum|launch__(block_size|grid_size|"
                  r"registers_per_thread|occupancy_limit_registers|waves_per_multiprocessor)|"
                  r"l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate\.pct|sass__inst_executed_local_(loads|stores)|"
                  r"sm__cycles_elapsed\.avg|sm__inst_executed_pipe_(lsu|alu|fma|fp64|xu)\.avg\.pct_of_peak_sustained_active|"
                  r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|"
                  r"smsp__inst_executed\.sum|smsp__thread_inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__warps_eligible\.avg\.per_cycle_active|smsp__sass_thread_inst_executed_op_.*_pred_on\.sum)$")

rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith("==")]
hdr, units, vals = rows[0], rows[1], rows[2]
w = csv.writer(sys.stdout, lineterminator="\n")
w.writerow(["metric", "unit", "value"])
for k, u, v in sorted(zip(hdr, units, vals)):
    if KEEP.match(k):
        w.writerow([k, u, v])
