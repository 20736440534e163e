"""ncu counters of the bench launch -> profiles/r02_ddp_<kind>_counters.json (read by bench.py's roofline object).

Run on the GPU box (one GPU; ncu replays the kernel once per metric group, so this takes a few times the kernel):

    python tools/capture_counters.py centroidal 16384        # or: srb 8192

Captures, for ONE launch of the solve kernel on the bench workload solved to convergence: DRAM bytes read / written,
executed DFMA thread-instructions, warp instructions, issue-slot and FP64-pipe utilisation, L2 hit rate — and stamps
them with the git revision of the sources, the batch and the mean DDP iteration count, so that bench.py can scale
them to its own launch and say which code they belong to.
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kind = sys.argv[1] if len(sys.argv) > 1 else "centroidal"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else (16384 if kind == "centroidal" else 8192)
METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "lts__t_sector_hit_rate.pct", "gpu__time_duration.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"]
runner = os.path.join(ROOT, "tools", "profile_solve.py") if kind == "centroidal" else os.path.join(ROOT, "tools", "profile_srb.py")
args = [str(batch), "0", "500"] if kind == "centroidal" else [str(batch), "500", "--once"]
cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", "regex:ddp_solve_kernel", "-c", "1", "--csv",
       sys.executable, runner] + args
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
r = subprocess.run(cmd, capture_output=True, text=True)
rows = [row for row in csv.reader(io.StringIO(r.stdout)) if len(row) > 10 and row[0].isdigit()]
if not rows:
    sys.stderr.write(r.stdout[-3000:] + r.stderr[-3000:])
    raise SystemExit("no ncu rows")
open(os.path.join(ROOT, "gpurun_out", f"r02_ddp_{kind}_counters_raw.csv"), "w").write(r.stdout)


def _num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return None


val = {row[-3]: _num(row[-1]) for row in rows}
unit = {row[-3]: row[-2] for row in rows}
mean_iters = None
for ln in r.stdout.splitlines():
    if ln.startswith("MEAN_ITERS"):
        mean_iters = float(ln.split()[1])
git = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip() or os.environ.get("CCC_GIT_REV", "unknown")
t = val["gpu__time_duration.sum"] * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit["gpu__time_duration.sum"].replace("second", "s").replace("nsecond", "ns"), 1e-9)
out = {
    "kernel": rows[0][4] if len(rows[0]) > 4 else "ddp_solve_kernel",
    "source": f"ncu metrics-only pass on one launch of the solve kernel, {batch} problems to convergence, sources at git {git} "
              f"(tools/capture_counters.py {kind} {batch})",
    "git": git, "batch": batch, "mean_ddp_iters": mean_iters,
    "dram_bytes_read": val["dram__bytes_read.sum"], "dram_bytes_write": val["dram__bytes_write.sum"],
    "dram_bytes_per_launch": val["dram__bytes_read.sum"] + val["dram__bytes_write.sum"],
    "dfma_thread_inst_per_launch": val["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"],
    "dadd_thread_inst_per_launch": val.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"),
    "dmul_thread_inst_per_launch": val.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"),
    "warp_inst_per_launch": val["smsp__inst_executed.sum"],
    "issue_active_pct": val.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "fp64_pipe_active_pct": val.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or val.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "l2_hit_rate_pct": val.get("lts__t_sector_hit_rate.pct"), "kernel_s_under_ncu": t,
}
path = os.path.join(ROOT, "gpurun_out", f"r02_ddp_{kind}_counters.json")
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
