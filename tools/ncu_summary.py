"""Key metrics of an ncu report:  python tools/ncu_summary.py <report.ncu-rep>"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak",
        "smsp__inst_executed.sum ", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_per_inst_issued", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_inst_executed_op_shared", "lts__t_sector_hit_rate", "sm__throughput.avg", "smsp__inst_executed.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(k.strip()) for k in KEYS):
        try:
            if float(v.replace(",", "")) == 0 and "stalled" in h:
                continue
        except ValueError:
            pass
        print(f"{h:86s} {u:10s} {v}")
