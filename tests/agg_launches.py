"""Profiling aid: aggregate the last N launches of an `ncu --metrics gpu__time_duration.sum --csv` log per kernel/grid."""
import collections, csv, re, sys
path, n = sys.argv[1], int(sys.argv[2])
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))[-n:]
def us(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
agg = collections.OrderedDict()
for r in rows:
    k = re.sub(r"\(.*", "", r["Kernel Name"])[:70] + " grid=" + r["Grid Size"] + " blk=" + r["Block Size"]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us(r)
tot = sum(a[1] for a in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  n={c:3d}  {k}")
print(f"step total {tot:.1f} us over {len(rows)} launches")
