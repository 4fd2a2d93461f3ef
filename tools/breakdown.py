#!/usr/bin/env python3
"""Aggregate the per-kernel table `bench.py --breakdown` prints to stderr: time by kernel family and the top shapes."""
import collections, re, sys
fam = collections.defaultdict(float); tot = 0.0; rows = []
for l in open(sys.argv[1]):
    m = re.match(r'(\S+?)\[(.*)\]\s+([\d.]+)\s+([\d.]+) ms\s+([\d.]+) GB/s', l)
    if m:
        name, tag, calls, ms, gbs = m.group(1), m.group(2), float(m.group(3)), float(m.group(4)), float(m.group(5))
        fam[name] += ms; tot += ms; rows.append((ms, name, tag, calls, gbs))
for k, v in sorted(fam.items(), key=lambda x: -x[1]):
    print(f"{k:28s} {v:7.2f} ms {100 * v / tot:5.1f}%")
print('total', round(tot, 2))
for r in sorted(rows, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{r[0]:6.3f} {r[1]}[{r[2]}] x{r[3]} {r[4]:.0f} GB/s")
