"""summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel launches, total and share of the summed kernel time"""
import csv, sys
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
tot = defaultdict(float); cnt = defaultdict(int)
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}.get(r[ui], 1e-6)
    name = r[ki].split("(")[0].replace("void ", "").replace("sacb::", "").replace("unnamed>::", "").replace("<unnamed>::", "").strip()
    tot[name] += v * scale; cnt[name] += 1
allms = sum(tot.values())
print(f"{'kernel':44s} {'launches':>8s} {'total ms':>12s} {'avg ms':>10s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:44]:44s} {cnt[k]:8d} {v:12.2f} {v / cnt[k]:10.2f} {100 * v / allms:6.1f}%")
print(f"{'sum':44s} {sum(cnt.values()):8d} {allms:12.2f}")
