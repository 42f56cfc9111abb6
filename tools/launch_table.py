"""Per-kernel shares of one steady-state EM iteration from an `ncu --metrics gpu__time_duration.sum --csv` launch list of
`bench.py --steps 2 --warmup 3 --no-cpu --no-e2e` (cold-cache, serialised: compare SHARES, not absolutes).

usage: python tools/launch_table.py <launches.csv> [iterations-in-the-list]
The list holds the set-up kernels and then (warm-up + timed + per-class profiling) identical iterations; the table is built
from the LAST iteration: the launches after the last but one `k_lmp` (the kernel that closes an iteration's M-step).
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
    names = [re.sub(r"^void |unnamed>::|\(.*$", "", r[4]).strip() for r in rows]
    ns = [float(r[-1]) for r in rows]
    ends = [i for i, n in enumerate(names) if n.startswith("k_lmp")]
    lo, hi = ends[-2] + 1, ends[-1] + 1
    agg = OrderedDict()
    for n, t in zip(names[lo:hi], ns[lo:hi]):
        c, s = agg.get(n, (0, 0.0))
        agg[n] = (c + 1, s + t)
    tot = sum(s for _, s in agg.values())
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for n, (c, s) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {n} | {c} | {s / 1e3:.1f} | {100 * s / tot:.1f} % |")
    print(f"| **total** | {hi - lo} | {tot / 1e3:.1f} | |")


if __name__ == "__main__":
    main()
