"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of each kernel (name + grid)."""
import collections, csv, re, sys
path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("b200sat::", "")
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
    v = v / 1000.0 if u == "ns" else (v * 1000.0 if u == "ms" else v)
    key = f"{name} grid={row.get('Grid Size','')}"
    agg[key][0] += 1; agg[key][1] += v; tot += v
print(f"total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{t / tot * 100:5.1f}%  n={n:4d}  avg={t / n:8.1f} us  {k[:120]}")
