"""Time per algorithm phase of the DDP solve kernel from an ncu report with warp-state samples.

    python tools/ncu_phases.py <report.ncu-rep> <lib.so that was profiled> [kernel-substring]

ncu attributes the samples of inlined helpers (dfma, shuffles, ld2 ...) to the helper's own source line, which
hides which part of the solver they belong to.  This tool re-attributes every SASS instruction to the last
*caller-level* source line seen before it in address order (a line of boxqp_warp.cuh, ddp_warp_core.cuh,
ddp_host.cuh or model_*.cuh; helper headers are skipped), using nvdisasm's line table of the same cubin, and
sums samples / executed instructions over the line ranges that make up each phase of the algorithm.
Code motion across phase boundaries blurs the split by a few instructions per boundary.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

CALLER_FILES = ("boxqp_warp.cuh", "ddp_warp_core.cuh", "ddp_host.cuh", "model_centroidal.cuh", "model_srb.cuh", "model_zmp.cuh")


def function_ranges(path, names):
    """{name: (first_line, last_line)} of the functions `names` in a source file (brace matching)."""
    src = open(path).read().split("\n")
    out = {}
    for name in names:
        for i, line in enumerate(src):
            if re.search(r"\b" + re.escape(name) + r"\s*\(", line) and not line.strip().endswith(";") and "CCC_DEV" in "".join(src[max(0, i - 1):i + 1]):
                depth, started = 0, False
                for j in range(i, len(src)):
                    depth += src[j].count("{") - src[j].count("}")
                    started = started or "{" in src[j]
                    if started and depth == 0:
                        out[name] = (i + 1, j + 1)
                        break
                break
    return out


def build_phase_map(root):
    csrc = os.path.join(root, "centroidalcontrolcollection_b200", "csrc")
    phases = []  # (file, lo, hi, phase)
    bq = function_ranges(os.path.join(csrc, "boxqp_warp.cuh"),
                         ["matvec32", "publish", "load_sym_row", "make_free_set", "load_compact_row", "llt_factor_compact",
                          "llt_back_compact", "llt_fwd_compact", "llt_solve_compactN", "boxqp_warp"])
    names = {"llt_factor_compact": "BoxQP: L D L' factorisation (+ fused forward substitution)",
             "llt_back_compact": "BoxQP: back substitution (1 rhs)", "llt_fwd_compact": "BoxQP: forward substitution (1 rhs, factor reused)",
             "llt_solve_compactN": "gains: K = -Quu_ff^-1 Qux_f (9 rhs)", "make_free_set": "BoxQP: free set + compact row gather",
             "load_compact_row": "BoxQP: free set + compact row gather", "load_sym_row": "BoxQP: row reload after factorisation",
             "matvec32": "BoxQP: matrix-vector products", "publish": "BoxQP: matrix-vector products"}
    for fn, (lo, hi) in bq.items():
        if fn != "boxqp_warp":
            phases.append(("boxqp_warp.cuh", lo, hi, names[fn]))
    if "boxqp_warp" in bq:
        phases.append(("boxqp_warp.cuh", bq["boxqp_warp"][0], bq["boxqp_warp"][1], "BoxQP: objective / Armijo / clamp logic"))
    core = function_ranges(os.path.join(csrc, "ddp_warp_core.cuh"), ["rollout", "step_and_cost", "terminal_cost", "backward_stage", "backward_pass", "solve", "quad", "storeX"])
    for fn, label in (("rollout", "forward passes (rollout + line search)"), ("step_and_cost", "forward passes (rollout + line search)"),
                      ("terminal_cost", "forward passes (rollout + line search)"), ("quad", "forward passes (rollout + line search)"),
                      ("storeX", "forward passes (rollout + line search)"), ("backward_pass", "backward pass: terminal + stage loop"),
                      ("solve", "DDP iteration control, outputs")):
        if fn in core:
            phases.append(("ddp_warp_core.cuh", core[fn][0], core[fn][1], label))
    if "backward_stage" in core:
        lo, hi = core["backward_stage"]
        src = open(os.path.join(csrc, "ddp_warp_core.cuh")).read().split("\n")
        cut1 = next(i + 1 for i in range(lo, hi) if "// gains" in src[i])
        cut2 = next(i + 1 for i in range(lo, hi) if "cost-to-go update" in src[i])
        phases.append(("ddp_warp_core.cuh", lo, cut1 - 1, "backward stage: derivatives, Q-functions, Quu assembly"))
        phases.append(("ddp_warp_core.cuh", cut1, cut2 - 1, "backward stage: BoxQP call, gain solve glue, gain store"))
        phases.append(("ddp_warp_core.cuh", cut2, hi, "backward stage: cost-to-go update (Quu K, Vx, Vxx)"))
    phases.append(("ddp_host.cuh", 1, 10 ** 6, "work queue (tickets, idle polling)"))
    for f in ("model_centroidal.cuh", "model_srb.cuh", "model_zmp.cuh"):
        mr = function_ranges(os.path.join(csrc, f), ["step", "lane_derivs", "init_Fx"])
        if "step" in mr:
            phases.append((f, mr["step"][0], mr["step"][1], "forward passes (rollout + line search)"))
        if "lane_derivs" in mr:
            phases.append((f, mr["lane_derivs"][0], mr["lane_derivs"][1], "backward stage: derivatives, Q-functions, Quu assembly"))
    return phases


def classify(phases, file, line):
    best = None
    for f, lo, hi, label in phases:
        if f == file and lo <= line <= hi:
            if best is None or (hi - lo) < best[0]:
                best = (hi - lo, label)
    return best[1] if best else f"other ({file})"


def main():
    rep, lib = sys.argv[1], sys.argv[2]
    pat = sys.argv[3] if len(sys.argv) > 3 else "CentroidalModelELi8ELi1ELb1"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    phases = build_phase_map(root)
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
    offs = {}  # offset -> phase
    for f in os.listdir(d):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        fn, ctx = None, ("ddp_host.cuh", 1)
        for line in out.splitlines():
            m = re.match(r"\.text\.(\S+):", line)
            if m:
                fn, ctx = m.group(1), ("ddp_host.cuh", 1)
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                base = os.path.basename(m.group(1))
                # small accessors / arithmetic helpers defined outside the phase functions keep the context
                if base in CALLER_FILES and not classify(phases, base, int(m.group(2))).startswith("other"):
                    ctx = (base, int(m.group(2)))
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", line)
            if m and fn and pat in fn:
                offs[int(m.group(1), 16)] = classify(phases, *ctx)
    csv_out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_out.splitlines()))
    hi = next(i for i, r in enumerate(rows[:10]) if r and r[0] == "Address")
    hdr = rows[hi]
    c_s, c_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
    base = None
    samp, inst = collections.Counter(), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        a = int(r[0], 16)
        base = a if base is None else base
        label = offs.get(a - base, "unmapped")
        samp[label] += int(r[c_s] or 0)
        inst[label] += int(r[c_i] or 0)
    ts, ti = sum(samp.values()), sum(inst.values())
    print(f"{'phase':72s} {'time (samples)':>14s} {'instructions':>13s}")
    for label, v in samp.most_common():
        print(f"{label:72s} {100 * v / ts:13.1f}% {100 * inst[label] / max(ti, 1):12.1f}%")
    print(f"total samples {ts}, warp instructions {ti}")


main()
