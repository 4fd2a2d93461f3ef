#!/usr/bin/env python3
"""Per kernel family: time in the step against its roofline floor, from a `bench.py --breakdown` table (stderr).
floor = max(algorithmic bytes / measured HBM peak, flops / sustained bf16 tensor peak) per launch shape, summed per family.
usage: excess_table.py profiles/live_breakdown_r02.txt [hbm_GBps] [tensor_TFLOPs]"""
import re
import sys

path = sys.argv[1]
hbm = float(sys.argv[2]) if len(sys.argv) > 2 else 6550.7
tc = float(sys.argv[3]) if len(sys.argv) > 3 else 1387.1
fam = {}
for line in open(path):
    m = re.match(r"^(\S+)\s+([\d.]+)\s+([\d.]+) ms\s+([\d.]+) GB/s\s+([\d.]+) TF/s", line)
    if not m:
        continue
    k, ms, gbs, tf = m.group(1), float(m.group(3)), float(m.group(4)), float(m.group(5))
    f = k.split("[")[0]
    if f == "gemm_tc":
        f = "gemm_tc(" + k.split("[")[1].split(":")[0] + ")"
    floor = max(ms * gbs / hbm, ms * tf / tc if f.startswith("gemm_tc") or f == "pw_bwd_fused" else 0.0)
    a = fam.setdefault(f, [0.0, 0.0])
    a[0] += ms
    a[1] += floor
tot, totf = sum(v[0] for v in fam.values()), sum(v[1] for v in fam.values())
print("| kernel family | ms/step | roofline floor (ms) | of floor | excess (ms) |")
print("|---|---:|---:|---:|---:|")
for f, (ms, fl) in sorted(fam.items(), key=lambda t: -(t[1][0] - t[1][1])):
    if ms - fl >= 0.15:
        print(f"| `{f}` | {ms:.2f} | {fl:.2f} | {fl / ms:.2f} | {ms - fl:.2f} |")
print(f"| all kernels | {tot:.2f} | {totf:.2f} | {totf / tot:.2f} | {tot - totf:.2f} |")
