"""Per-source-line sample summary of an ncu report:  python tools/ncu_lines.py <report.ncu-rep> [top]
Prints, for the top source lines by stall samples, total samples, instructions executed and the
dominant stall reasons (needs --import-source on at capture and -lineinfo at compile time)."""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = next(i for i, r in enumerate(rows[:20]) if "Line No" in r)
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    c_samp = hdr.index("# Samples")
    c_inst = hdr.index("Instructions Executed")
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    lines = []
    file_name = ""
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        if r[0] == "":
            continue
        try:
            samp = int(r[c_samp])
        except ValueError:
            continue
        inst = int(r[c_inst]) if r[c_inst].isdigit() else 0
        st = sorted(((int(r[i]) if r[i].isdigit() else 0, h) for h, i in stall_cols), reverse=True)[:3]
        lines.append((samp, inst, r[0], r[1].strip()[:90], st))
    tot = sum(l[0] for l in lines)
    toti = sum(l[1] for l in lines)
    print(f"total samples {tot}, instructions executed {toti}")
    agg = {}
    for l in lines:
        for v, h in l[4]:
            pass
    for samp, inst, ln, src, st in sorted(lines, reverse=True)[:top]:
        s = " ".join(f"{h[6:]}={v}" for v, h in st if v)
        print(f"{100*samp/tot:5.1f}% samp  {100*inst/max(toti,1):5.1f}% inst  L{ln:>4} {src:90s} {s}")


main()
