"""Attribute SASS instruction counts to source lines:  python tools/sass_lines.py <lib.so> <kernel-substring> [top]
(nvdisasm -g on the embedded cubin; needs -lineinfo at compile time)."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    d = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
    cnt, per_file = collections.Counter(), collections.Counter()
    total = 0
    for f in os.listdir(d):
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
            if fn and pat in fn and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
                cnt[cur] += 1
                total += 1
    print("total SASS instructions:", total, f"({total * 16 / 1024:.0f} KiB)")
    for k, v in cnt.most_common(top):
        print(f"{v:6d}  {k}")


main()
