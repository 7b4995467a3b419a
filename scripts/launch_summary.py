"""aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python scripts/launch_summary.py launches.csv"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("void ", "")
    name = re.sub(r"<.*", "", name) if name.startswith("at::") else name
    agg[name][0] += 1
    agg[name][1] += float(r[-1]) / 1e3
total = sum(v[1] for v in agg.values())
print(f"# launch list: {len(rows)} launches, {total:.1f} us total (ncu-serialised, cold-cache: compare SHARES, not absolutes)\n")
print("| kernel | launches | total us | share |\n|---|---|---|---|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {t:.1f} | {100 * t / total:.1f} % |")
