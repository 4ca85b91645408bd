"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel family."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot = defaultdict(float)
cnt = defaultdict(int)
for row in r:
    if row[mi] != "gpu__time_duration.sum":
        continue
    v = float(row[vi].replace(",", ""))
    u = row[ui]
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1, "second": 1e9}.get(u, 1)
    name = row[ki].replace("(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"\(.*", "", name).replace("at::native::", "")[:90]
    tot[name] += ns
    cnt[name] += 1
allns = sum(tot.values())
print(f"total {allns/1e6:.3f} ms over {sum(cnt.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v/1e6:9.3f} ms {100*v/allns:5.1f}%  x{cnt[k]:<5d} {k}")
