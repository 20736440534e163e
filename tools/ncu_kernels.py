"""Per-kernel key metrics of an ncu report (one block per profiled launch):  python tools/ncu_kernels.py <report.ncu-rep>"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit",
        "sm__warps_active.avg.pct_of_peak", "smsp__issue_active.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "sm__pipe_tensor",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_per_inst_issued",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate", "sm__throughput.avg", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
name_col = hdr.index("Kernel Name") if "Kernel Name" in hdr else 4
for vals in rows[2:]:
    if len(vals) != len(hdr):
        continue
    print(f"=== {vals[name_col][:150]}")
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(k.strip()) for k in KEYS):
            try:
                if float(v.replace(",", "")) == 0 and "stalled" in h:
                    continue
            except ValueError:
                pass
            print(f"  {h:84s} {u:10s} {v}")
