"""Code map of a kernel from an ncu report: where in the instruction stream the time goes, and how much of it is
instruction fetch.

    python tools/ncu_code_map.py <report.ncu-rep> <lib.so that was profiled> <mangled-kernel-substring> [bucket bytes]

The kernel's SASS is cut into buckets of `bucket bytes` (default 2 KB = 128 instructions) in address order; for every bucket
that was executed: executions per instruction (tight loops stand out from once-per-stage straight-line code), share of the
warp-state samples, the `no_instruction` share of those samples (instruction-cache misses / fetch stalls), and the source
lines (caller-level files only) most of its instructions come from.  Needs -lineinfo at compile time and
--import-source on at capture; the library must be the build that was profiled.
Found with it (profiles/r02_summary.md §5.2): the 12-state DDP kernel spent a quarter of its stall samples fetching 44 KB of
once-per-stage straight-line code.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

HELPERS = ("warp_ctx.cuh", "sm_30_intrinsics.hpp", "sm_32_intrinsics.hpp", "device_atomic_functions.hpp", "device_functions.h",
           "math_functions.hpp")


def line_table(lib, pat):
    """{offset in the kernel: (file, line)} from nvdisasm's line info of the cubins embedded in `lib`."""
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
    table = {}
    for f in sorted(os.listdir(d)):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        cur, fn = None, None
        for line in out.splitlines():
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\.text\.(\S+):", line)
            if m:
                fn = m.group(1)
                continue
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", line)
            if fn and pat in fn and m:
                table.setdefault(int(m.group(1), 16), cur)
    return table


def main():
    rep, lib, pat = sys.argv[1], sys.argv[2], sys.argv[3]
    bucket = int(sys.argv[4], 0) if len(sys.argv) > 4 else 0x800
    table = line_table(lib, pat)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows[:20]) if "Address" in r)
    col = {h: i for i, h in enumerate(rows[hi])}
    data = [r for r in rows[hi + 1:] if len(r) > col["stall_no_inst"] and r[0].startswith("0x")]
    base = int(data[0][0], 16)
    ie, ni, ns = col["Instructions Executed"], col["stall_no_inst"], col["# Samples"]
    buckets = collections.OrderedDict()
    for r in data:
        a = int(r[0], 16) - base
        d = buckets.setdefault(a // bucket, [0, 0, 0, 0, collections.Counter()])
        d[0] += 1
        d[1] += int(r[ie])
        d[2] += int(r[ns])
        d[3] += int(r[ni])
        ln = table.get(a)
        if ln and ln[0] not in HELPERS and not (ln[0] == "boxqp_warp.cuh" and ln[1] < 60):
            d[4][(ln[0][:16], ln[1] // 10 * 10)] += 1
    tot = sum(d[2] for d in buckets.values())
    hot = sum(d[0] for d in buckets.values() if d[1] > 0)
    print(f"{len(data)} instructions ({len(data) * 16 // 1024} KiB), buckets executed at least once hold {hot} ({hot * 16 // 1024} KiB); "
          f"{tot} samples, {100 * sum(d[3] for d in buckets.values()) / max(tot, 1):.1f} % of them no_instruction")
    for b, d in buckets.items():
        if d[1] == 0:
            continue
        top = ", ".join(f"{k[0]}:{k[1]}" for k, _ in d[4].most_common(3))
        print(f"{b * bucket:6x} exec/instr {d[1] / d[0] / 1e6:7.2f}M samples {100 * d[2] / tot:5.2f}% no_inst {100 * d[3] / max(d[2], 1):4.0f}%  {top}")


main()
