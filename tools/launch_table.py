"""Markdown table per kernel from an ncu launch list
(`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file X.csv ...`).
usage: launch_table.py X.csv [peak_gbs=6543.4]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6543.4
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
ix = {h: i for i, h in enumerate(rows[0])}
per = collections.defaultdict(dict)
name = {}
for r in rows[1:]:
    per[r[ix["ID"]]][r[ix["Metric Name"]]] = (float(r[ix["Metric Value"]].replace(",", "")), r[ix["Metric Unit"]])
    name[r[ix["ID"]]] = r[ix["Kernel Name"]]
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
for i, m in per.items():
    k = name[i]
    short = re.sub(r"\(.*", "", k).replace("zk::<unnamed>::", "").replace("void ", "")
    if "zk::" not in k:
        short = "(torch: fills / copies / checks)"
    a = agg[short]
    a[0] += 1
    t, u = m["gpu__time_duration.sum"]
    a[1] += t * scale[u]
    for j, key in ((2, "dram__bytes_read.sum"), (3, "dram__bytes_write.sum")):
        if key in m:
            a[j] += m[key][0] * scale[m[key][1]]
total = sum(a[1] for a in agg.values())
print("| kernel | launches | total ms | share | DRAM read MB / launch | DRAM write MB / launch | DRAM GB/s | of peak |")
print("|---|---:|---:|---:|---:|---:|---:|---:|")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (a[2] + a[3]) / (a[1] * 1e-3) / 1e9 if a[1] else 0
    print("| `%s` | %d | %.3f | %.1f %% | %.1f | %.1f | %.0f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / total, a[2] / a[0] / 1e6, a[3] / a[0] / 1e6, gbs, 100 * gbs / peak))
print("\ntotal %.3f ms over %d launches" % (total, sum(a[0] for a in agg.values())))
