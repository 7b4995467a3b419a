"""Per-opcode dynamic instruction counts and stall samples from an ncu report's source page:
    python scripts/sass_hot.py gpurun_out/prof_x.ncu-rep [kernel-substring]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, cs, cx = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
ops = collections.Counter()
samples = collections.Counter()
tot = 0
top = []
for r in rows[hi + 1:]:
    if len(r) <= cx:
        continue
    src = r[ci].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    if not m:
        continue
    op = m.group(2)
    n = int(r[cx] or 0)
    s = int(r[cs] or 0)
    ops[op] += n
    samples[op] += s
    tot += n
    top.append((s, n, src))
print(f"total warp instructions {tot}")
print("| opcode | executed | share | stall samples |\n|---|---|---|---|")
for op, n in ops.most_common(28):
    print(f"| {op} | {n} | {100*n/tot:.1f} % | {samples[op]} |")
print("\nhottest lines by samples:")
for s, n, src in sorted(top, reverse=True)[:25]:
    print(f"{s:7d} {n:10d}  {src[:110]}")
