"""summarise an ncu gpu__time_duration launch list: python tools/launch_report.py file.csv start_kernel_substr [--list]"""
import collections
import csv
import re
import sys

fn, key = sys.argv[1], sys.argv[2]
rows = []
with open(fn) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
for row in csv.DictReader(lines):
    if row.get("Metric Name") == "gpu__time_duration.sum":
        rows.append((re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:56], float(row["Metric Value"].replace(",", "")),
                     row["Grid Size"]))
starts = [i for i, r in enumerate(rows) if key in r[0]]
seg = rows[starts[-1]:] if len(starts) < 2 or "--last" in sys.argv else rows[starts[-2]:starts[-1]]
tot, cnt = collections.defaultdict(float), collections.Counter()
for n, v, g in seg:
    tot[n] += v
    cnt[n] += 1
T = sum(tot.values())
print(f"# segment: {len(seg)} launches, {T / 1e6:.3f} ms of kernel time")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"# {100 * v / T:5.1f}%  {v / 1e6:7.3f} ms  n={cnt[k]:3d}  {k}")
if "--list" in sys.argv:
    for n, v, g in seg:
        print(f"{v / 1e3:9.1f} us  {g:>16s}  {n}")
