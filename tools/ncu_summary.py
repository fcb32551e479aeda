"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time per frame."""
import collections, csv, re, sys

path, frames = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rows = list(csv.DictReader(lines))
def us(r):
    v = float(r["Metric Value"].replace(",", "")); u = r["Metric Unit"]
    return v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
agg, tot = collections.OrderedDict(), 0.0
for r in rows:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("vls::<unnamed>::", "").replace("void ", "")[:60]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += us(r); tot += us(r)
print(f"{len(rows)} launches, {tot / frames:.1f} us/frame over {frames:g} frame(s) (cold-cache, serialised: compare shares)")
for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / frames:9.1f} us/frame {c / frames:6.1f} launches/frame {100 * t / tot:5.1f}%  {n}")
