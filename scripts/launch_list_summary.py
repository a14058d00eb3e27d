"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (last step only when --last-step N is
given: the final N launches).  Usage: launch_list_summary.py <launches.csv> <out.csv> [launches_per_step]"""
import csv, sys
from collections import defaultdict

src, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
launches = [(r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("sonic::", ""), float(r[col["Metric Value"]].replace(",", "")))
            for r in rows[1:] if r[col["Metric Name"]] == "gpu__time_duration.sum"]
if len(sys.argv) > 3:
    launches = launches[-int(sys.argv[3]):]
tot = defaultdict(lambda: [0, 0.0])
for k, ns in launches:
    tot[k][0] += 1
    tot[k][1] += ns
total = sum(v[1] for v in tot.values())
with open(out, "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "total_ms", "share"])
    for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        w.writerow([k, n, round(ns / 1e6, 3), round(ns / total, 4)])
    w.writerow(["TOTAL", sum(v[0] for v in tot.values()), round(total / 1e6, 3), 1.0])
print(open(out).read())
