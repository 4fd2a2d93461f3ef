#!/usr/bin/env python3
"""Kernel-family shares of the training step in two views that must agree: bench.py's live CUDA-event table (no profiler) and
the ncu launch list of the same command (cold-cache, serialised).  usage: share_table.py live_breakdown.txt launches.csv"""
import collections, csv, re, sys

FAMILY = [  # (family, live-table prefix regex, ncu kernel-name regex)
    ("gemm_tc (nt + convt)", r"gemm_tc\[(nt|convt):", r"gemm_tc_nt_kernel"),
    ("gemm_tc (wgrad)", r"gemm_tc\[wgrad:", r"gemm_tc_wgrad_kernel"),
    ("pw_bwd_fused", r"pw_bwd_fused\[", r"pw_bwd_fused_kernel"),
    ("dwconv3x3_bwd", r"dwconv3x3_bwd\[", r"dwconv3x3_bwd_strip_kernel"),
    ("dwconv3x3_fwd", r"dwconv3x3_fwd\[", r"dwconv3x3_strip_kernel|dwconv3x3_vec8|dwconv3x3_scalar"),
    ("maxpool2x2_bwd", r"maxpool2x2_bwd\[", r"maxpool_bwd_kernel"),
    ("bn_act", r"bn_act\[", r"bn_act_kernel"),
    ("bn_bwd_apply", r"bn_bwd_apply\[", r"bn_bwd_apply_kernel"),
    ("bn_bwd_reduce", r"bn_bwd_reduce\[", r"bn_bwd_reduce_kernel"),
    ("head_bwd", r"head_bwd\[", r"head1?_bwd"),
    ("head_fwd", r"head_fwd\[", r"head1?_fwd"),
    ("stem_bwd_folded", r"stem_bwd_folded\[", r"stem_bwd_folded_kernel"),
    ("stem_fwd", r"stem_fwd\[", r"stem_dw_kernel|stem_pw_kernel|stem_fwd_kernel"),
    ("convt_bwd_gather", r"convt_bwd_gather\[", r"convt_bwd_gather_kernel"),
]
live = collections.defaultdict(float); live_tot = 0.0
for line in open(sys.argv[1]):
    m = re.match(r"(\S+)\s+([\d.]+)\s+([\d.]+) ms", line)
    if not m:
        continue
    ms = float(m.group(3)); live_tot += ms
    for fam, lre, _ in FAMILY:
        if re.match(lre, m.group(1)):
            live[fam] += ms
            break
ncu = collections.defaultdict(float); ncu_tot = 0.0
rows = list(csv.DictReader([l for l in open(sys.argv[2]) if l.startswith('"')]))
for r in rows:
    try:
        v = float(r["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = r.get("Metric Unit", "ns")
    ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
    ncu_tot += ms
    for fam, _, nre in FAMILY:
        if re.search(nre, r["Kernel Name"]):
            ncu[fam] += ms
            break
steps = max(1, sum(1 for r in rows if "adamw_kernel" in r["Kernel Name"]))       # one AdamW launch per training step
print(f"training steps under ncu: {steps}\n")
print("| kernel family | live CUDA events: ms/step (share) | ncu launch list: ms/step (share) |\n|---|---:|---:|")
for fam, _, _ in sorted(FAMILY, key=lambda f: -live[f[0]]):
    print(f"| `{fam}` | {live[fam]:.2f} ({100 * live[fam] / live_tot:.1f} %) | {ncu[fam] / steps:.2f} ({100 * ncu[fam] / ncu_tot:.1f} %) |")
print(f"| total | {live_tot:.2f} | {ncu_tot / steps:.2f} |")
