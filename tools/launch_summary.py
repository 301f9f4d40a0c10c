"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv).
Usage: python tools/launch_summary.py gpurun_out/x_launches.csv > profiles/x_launches_summary.txt"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.defaultdict(float), collections.Counter()
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki])[:200]
    tot[name] += float(r[vi].replace(",", "")) * scale[r[ui]]
    cnt[name] += 1
allms = sum(tot.values())
print("# ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 1500 python bench.py --steps 1 --warmup 1 "
      "(serialised, cold-cache: compare SHARES)")
print(f"# {sum(cnt.values())} launches, {allms:.1f} ms total")
for n, t in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{t:10.2f} ms  {100 * t / allms:5.1f}%  {cnt[n]:5d} launches  {n}")
