#!/usr/bin/env python
"""Summarise the LAST training step of an `ncu --metrics gpu__time_duration.sum --csv` log of bench.py
(the step is delimited by the first kernel of the forward pass: stem_im2col / lstm_xproj / layer-norm)."""
import collections, csv, sys
path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "stem_im2col"
rows, hdr = [], None
with open(path) as f:
    for row in csv.reader(f):
        if hdr is None:
            if "Kernel Name" in row:
                hdr = row
            continue
        rows.append(row)
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
names = [r[ik] for r in rows]
vals = [float(r[iv].replace(",", "")) for r in rows]
idx = [i for i, n in enumerate(names) if marker in n]
start = idx[-1] if idx else 0
agg = collections.defaultdict(lambda: [0, 0.0])
for n, v in zip(names[start:], vals[start:]):
    k = n.split("(")[0][:100]
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v for _, v in agg.values())
print(f"last step: {len(names) - start} launches, {tot / 1e6:.3f} ms (ncu-serialised, cold caches: compare shares)")
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{v / 1e6:8.3f} ms {100 * v / tot:5.1f}%  x{c:<4d} {k}")
