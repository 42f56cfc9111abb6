"""Per-source-line profile from an .ncu-rep captured with --import-source on (no GPU needed).

`ncu --page source --csv` exports per-SASS-instruction counters without line numbers; `nvdisasm
--print-line-info-inline` on the cubin of the same build gives the line of every instruction (with its inlining
chain).  The two listings have the same instruction order, so they are joined by index and aggregated by the
outermost line inside the given source file.

usage: python tools/ncu_lines.py <rep> <object.o> <kernel-regex> <source-file-basename> [launch-skip]
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os
from collections import defaultdict


def sass_rows(rep, kregex, skip):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex,
                          "--launch-skip", str(skip), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    h = rows[1]
    return name, [dict(zip(h, r)) for r in rows[2:] if len(r) == len(h)]


def line_map(obj, mangled_hint, srcbase):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cub = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    txt = subprocess.run(["nvdisasm", "--print-line-info-inline", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
    funcs = {}
    cur, pend = None, []
    for ln in txt.splitlines():
        m = re.match(r"^\.text\.(\S+):$", ln)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            pend = []
            continue
        if cur is None:
            continue
        if "//## File" in ln:
            pend.append(ln)
            continue
        m = re.match(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            if pend:
                # outermost frame inside srcbase: last '//## File "<srcbase>", line N' line of the group
                line = None
                inner = None
                for p in pend:
                    mm = re.search(r'File "([^"]+)", line (\d+)', p)
                    if mm and inner is None:
                        inner = (os.path.basename(mm.group(1)), int(mm.group(2)))
                    for mm2 in re.finditer(r'"([^"]+)", line (\d+)', p):
                        if os.path.basename(mm2.group(1)) == srcbase:
                            line = int(mm2.group(2))
                last = (line, inner)
                pend = []
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip(), last))
    for k, v in funcs.items():
        if mangled_hint in k:
            return v
    raise SystemExit("kernel not found in cubin: " + mangled_hint + " candidates: " + ", ".join(funcs))


if __name__ == "__main__":
    rep, obj, kregex, srcbase = sys.argv[1:5]
    skip = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    name, rows = sass_rows(rep, kregex, skip)
    m = re.search(r"k_\w+<[^>]*>", name)
    # build the mangled hint from template ints, e.g. k_mcpass<(int)3, (int)3> -> k_mcpassILi3ELi3E
    base = re.search(r"(k_\w+)<", name).group(1)
    ints = re.findall(r"\(int\)(\d+)", name)
    hint = base + "I" + "".join(f"Li{v}E" for v in ints)
    lm = line_map(obj, hint, srcbase)
    if len(lm) != len(rows):
        print(f"warning: {len(lm)} instructions in the cubin vs {len(rows)} in the report", file=sys.stderr)
    agg = defaultdict(lambda: [0, 0, 0, 0])
    tot = [0, 0, 0]
    inner_agg = defaultdict(lambda: [0, 0])
    for (off, ins, (line, inner)), r in zip(lm, rows):
        ie = int(float(r.get("Instructions Executed", "0") or 0))
        te = int(float(r.get("Thread Instructions Executed", "0") or 0))
        ss = int(float(r.get("Warp Stall Sampling (All Samples)", "0") or 0))
        a = agg[line]
        a[0] += ie
        a[1] += te
        a[2] += ss
        a[3] += 1
        tot[0] += ie
        tot[1] += te
        tot[2] += ss
        ia = inner_agg[inner]
        ia[0] += ie
        ia[1] += ss
    src = {}
    for root in ("simples_b200/csrc",):
        p = os.path.join(root, srcbase)
        if os.path.exists(p):
            src = {i + 1: l.rstrip() for i, l in enumerate(open(p))}
    print(f"# {name}\n# total warp instructions {tot[0]:.4g}, thread instructions {tot[1]:.4g}, samples {tot[2]}")
    print("# line | warp-inst % | samples % | lanes | sass | source")
    for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if a[0] < 0.003 * tot[0] and a[2] < 0.003 * tot[2]:
            continue
        lanes = a[1] / a[0] if a[0] else 0
        print(f"{str(line):>5} | {100.0 * a[0] / tot[0]:6.2f} | {100.0 * a[2] / max(tot[2], 1):6.2f} | {lanes:5.1f} | {a[3]:4d} | {src.get(line, '')[:110]}")
    print("# by innermost inlined location (top 25)")
    for inner, a in sorted(inner_agg.items(), key=lambda kv: -kv[1][0])[:25]:
        print(f"{str(inner):>40} | {100.0 * a[0] / tot[0]:6.2f} | {100.0 * a[1] / max(tot[2], 1):6.2f}")
