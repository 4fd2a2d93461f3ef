#!/usr/bin/env python3
"""Summarise gpurun_out/*.ncu-rep and launches.csv into tracked files under profiles/ (run here, no GPU needed)."""
import csv, io, json, os, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("UNET_PROFILES_OUT") or os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
summary = {}
for f in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
    if not f.endswith(".ncu-rep"):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", f), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                rec[h] = f"{v} {u}".strip()
        summary[f[:-8]] = rec
json.dump(summary, open(os.path.join(OUT, f"ncu_full_{tag}.json"), "w"), indent=1)
# DRAM bytes (read + write) per launch of each captured kernel, keyed by bench.py's per-kernel table tag (roofline.traffic)
BENCH_KEY = {"prof_dw_fwd128": "dwconv3x3_fwd[64x512x512x128]", "prof_dw_fwd_aff": "dwconv3x3_fwd[64x512x512x64]",
             "prof_gemm64": "gemm_tc[nt:16777216x64x64:e3]", "prof_dw_bwd_aff": "dwconv3x3_bwd[64x512x512x64+mask]",
             "prof_dw_bwd_mask": "dwconv3x3_bwd[64x512x512x64+mask,noaffine]",
             "prof_pw_bwd_fused64": "pw_bwd_fused[16777216x64x128]", "prof_pw_bwd_fused128": "pw_bwd_fused[16777216x128x128]",
             "prof_gemm_fold_dgrad": "gemm_tc[nt:16777216x128x128:e1]", "prof_gemm_fold_wgrad": "gemm_tc[wgrad:128x128x16777216:e0]",
             "prof_bn_bwd_apply": "bn_bwd_apply[16777216x64]", "prof_fused64": "sepconv_fused_fwd[64x512x512x64->64]",
             "prof_bn_act": "bn_act[64x512x512x64]"}
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
def _bytes(v):
    num, unit = v.split()[0], (v.split() + ["byte"])[1]
    return float(num.replace(",", "")) * UNIT.get(unit, 1)
traffic = {}
for k, rec in summary.items():
    if k in BENCH_KEY and "dram__bytes_read.sum" in rec and "dram__bytes_write.sum" in rec:
        traffic[BENCH_KEY[k]] = int(_bytes(rec["dram__bytes_read.sum"]) + _bytes(rec["dram__bytes_write.sum"]))
if traffic:
    json.dump(traffic, open(os.path.join(OUT, "ncu_traffic.json"), "w"), indent=1)
with open(os.path.join(OUT, f"ncu_full_{tag}.md"), "w") as o:
    o.write(f"# ncu --set full --clock-control none summaries ({tag}); source: tools/ncu_capture.sh, shapes = train512 level 0 (batch 64, 512x512)\n\n")
    for k, rec in summary.items():
        o.write(f"## {k}: `{rec['kernel'][:110]}`\n\n| metric | value |\n|---|---|\n")
        for h in WANT:
            if h in rec:
                o.write(f"| {h} | {rec[h]} |\n")
        o.write("\n")
# launch list: per-kernel totals and share of the step
path = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = r.get("Metric Unit", "ns")
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        name = r["Kernel Name"].split("(")[0]
        name = name.replace("unet::", "")
        tot[name][0] += 1; tot[name][1] += ms
    total = sum(v[1] for v in tot.values())
    with open(os.path.join(OUT, f"ncu_launches_{tag}.md"), "w") as o:
        o.write(f"# ncu launch list ({tag}): `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-infer` under\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` — every launch of 5 training steps (train512, batch 64).\n"
                "Per-launch times are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event table, not absolutes.\n\n")
        o.write(f"launches: {sum(v[0] for v in tot.values())}, total kernel time {total:.1f} ms\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            o.write(f"| `{k[:90]}` | {n} | {ms:.2f} | {100 * ms / total:.1f}% |\n")
    import shutil
    shutil.copy(path, os.path.join(OUT, f"ncu_launches_{tag}.csv"))
print("written", os.listdir(OUT))
