"""Static SASS listing of one kernel annotated with the outermost source line of a given file (no GPU needed).

usage: python tools/sass_lines.py <object.o> <kernel-substring> <source-basename> [--list]
Prints the static instruction count per source line (and, with --list, the annotated listing) -- the offline proxy used to
check instruction-trimming changes of k_sweep before spending GPU time.
"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter


def annotate(obj, ksub, srcbase):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    cur, pend, out, last = None, [], [], (None, ("?", 0))
    for ln in txt.splitlines():
        m = re.match(r"^\.text\.(\S+):$", ln)
        if m:
            cur, pend = m.group(1), []
            continue
        if cur is None or ksub not in cur:
            continue
        if "//## File" in ln:
            pend.append(ln)
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if pend:
                line, inner = None, None
                for p in pend:
                    for mm in re.finditer(r'"([^"]+)", line (\d+)', p):
                        b = os.path.basename(mm.group(1))
                        if inner is None:
                            inner = (b, int(mm.group(2)))
                        if b == srcbase:
                            line = int(mm.group(2))
                last = (line, inner or ("?", 0))
            pend = []
            out.append((m.group(1), last[0], last[1], m.group(2).strip()))
        elif re.match(r"^\s*\.L_", ln):
            out.append((None, None, None, ln.strip()))
    return out


if __name__ == "__main__":
    obj, ksub, srcbase = sys.argv[1:4]
    rows = annotate(obj, ksub, srcbase)
    if "--list" in sys.argv:
        for a, l, inner, ins in rows:
            print(ins if a is None else f"{a} {l} {inner[0]}:{inner[1]} | {ins}")
    else:
        c = Counter(l for a, l, _, _ in rows if a is not None)
        tot = sum(c.values())
        print(f"# {tot} instructions")
        for l in sorted(k for k in c if k is not None):
            print(f"{l:5d} {c[l]:5d}")
