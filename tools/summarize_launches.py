"""Aggregate an ncu launch list (--csv --metrics gpu__time_duration.sum) by kernel: count, total us, share."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r]
    if not hi:
        raise SystemExit("no ncu csv header in " + path)
    h = rows[hi[0]]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hi[0] + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = re.sub(r"\(.*", "", r[ki]).replace("void ", "")[:80]
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"{path}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.1f} us (serialised, cold-cache ncu times)")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {k:80s} n={v[0]:5d} {v[1] / 1e3:10.1f} us {100 * v[1] / tot:5.1f}%")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
