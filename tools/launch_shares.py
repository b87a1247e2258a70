"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: launches, total
device time and share.  Per-launch times under ncu are cold-cache and serialised, so compare SHARES
with bench.py's "kernels" table, not absolutes.   usage: launch_shares.py launches.csv [out.txt]"""
import csv, re, sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[12] == "gpu__time_duration.sum"]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(int\)|\(bool\)", "", r[4])
    name = re.sub(r"\(.*$", "", name).replace("void ", "").replace("bc::", "")
    a = agg[name]
    a[0] += 1
    a[1] += float(r[14]) / 1e3
tot = sum(v[1] for v in agg.values())
lines = [f"{len(rows)} launches, {tot / 1e3:.3f} ms of kernel time under ncu (cold-cache, serialised)",
         f"{'kernel':72s} {'n':>5s} {'us total':>11s} {'us/launch':>10s} {'share':>6s}"]
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"{k:72s} {n:5d} {us:11.1f} {us / n:10.1f} {us / tot:6.3f}")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write(out + "\n")
