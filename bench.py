#!/usr/bin/env python3
"""bench.py — the hot path on synthetic data: U-Net (reference model/u_net.py) Dice-loss training or inference, img/s.

  python bench.py --gpus N --steps K --warmup W            # this framework (sm_100a kernels via the C-ABI)
  python bench.py --impl reference --gpus N --steps K ...  # the CPU arm: torch-CPU port of the reference's TF path

Default workload = BASELINE.json configs[2] (the 512x512 configuration the metric is quoted on): U_NET((512,512,3)),
binary mask, Dice loss + Keras-form AdamW, dropout 0.2, batch 64 per GPU, bf16 activations / fp32 accumulation,
data-parallel over N GPUs (weak scaling; one NCCL gradient all-reduce per step, overlapped with backward).
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (mode, H, W, classes, batch per GPU, description)
    "train512": ("train", 512, 512, 1, 64, "BASELINE configs[2]: U-Net 512x512 Dice-loss training, batch 64/GPU"),
    "train256": ("train", 256, 256, 1, 32, "BASELINE configs[1]: U-Net 256x256 Dice-loss training, batch 32"),
    "train512c8": ("train", 512, 512, 8, 32, "BASELINE configs[4]: 8-class softmax U-Net 512x512 training, batch 32/GPU"),
    "infer256": ("infer", 256, 256, 1, 8, "BASELINE configs[0]: U-Net 256x256 inference, batch 8"),
    "infer512": ("infer", 512, 512, 1, 64, "U-Net 512x512 inference, batch 64/GPU"),
    "infer1024": ("infer", 1024, 1024, 1, 16, "BASELINE configs[3]: U-Net 1024x1024 inference, batch 16/GPU"),
}
RIDGE_FLOP_PER_BYTE = 212.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="train512", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print the per-kernel table to stderr")
    ap.add_argument("--no-graphs", action="store_true", help="launch every step eagerly (no CUDA-graph replay)")
    ap.add_argument("--no-infer", action="store_true", help="training workloads: skip the inference record (`infer`) of the same shape")
    ap.add_argument("--check", action="store_true", help="N > 1: data-parallel correctness leg (gradient == mean of shard gradients; "
                    "weights identical across ranks after steps) before the timed runs")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ====================================================================================================== CPU arm
def cpu_train_sample(H, W, classes, batch, steps, warmup=1):
    """The oracle's torch-CPU port of the reference training step (oracle/torch_ref.py), all host threads."""
    import numpy as np
    import torch
    from oracle import torch_ref as TR
    from oracle import unet_ref as R
    specs = R.layer_specs((H, W, 3), classes, 0.2, True)
    P = R.init_params(specs, seed=2301)
    x, y = R.synthetic_batch(batch, H, W, 3, classes, seed=2301)
    Pt = TR.to_torch(P, dtype=torch.float32, requires_grad=True)
    train = [v for v in Pt.values() if v.requires_grad]
    m = [torch.zeros_like(v) for v in train]
    v2 = [torch.zeros_like(v) for v in train]
    xt, yt = torch.tensor(x), torch.tensor(y)
    times = []
    for t in range(1, warmup + steps + 1):
        t0 = time.perf_counter()
        probs = TR.forward(Pt, xt, classes, 0.2, True, training=True, drop_seeds=None)
        loss = 1.0 - TR.dice_coef(yt, probs)
        grads = torch.autograd.grad(loss, train)
        with torch.no_grad():      # Keras-form AdamW (train.py:226)
            a = 2e-3 * (1 - 0.999 ** t) ** 0.5 / (1 - 0.9 ** t)
            for w_, g_, m_, v_ in zip(train, grads, m, v2):
                w_.mul_(1 - 2e-3 * 1e-4)
                m_.mul_(0.9).add_(g_, alpha=0.1)
                v_.mul_(0.999).addcmul_(g_, g_, value=0.001)
                w_.addcdiv_(m_, v_.sqrt().add_(1e-7), value=-a)
        float(loss)
        if t > warmup:
            times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def cpu_infer_sample(H, W, classes, batch, steps, warmup=1):
    import torch
    from oracle import torch_ref as TR
    from oracle import unet_ref as R
    specs = R.layer_specs((H, W, 3), classes, 0.2, True)
    Pt = TR.to_torch(R.init_params(specs, seed=2301), dtype=torch.float32)
    x, _ = R.synthetic_batch(batch, H, W, 3, classes, seed=2301)
    xt = torch.tensor(x)
    times = []
    with torch.no_grad():
        for t in range(warmup + steps):
            t0 = time.perf_counter()
            TR.forward(Pt, xt, classes, 0.2, True, training=False).sum().item()
            if t >= warmup:
                times.append(time.perf_counter() - t0)
    return batch * len(times) / sum(times), sum(times) / len(times)


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_sample_size(mode, H, W):
    """Bounded sample: a few images so one CPU step is seconds, not minutes."""
    px = H * W
    if mode == "train":
        return max(1, min(8, (2 * 512 * 512) // px))
    return max(1, min(8, (8 * 256 * 256) // px))


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    mode, H, W, classes, _, desc = WORKLOADS[args.workload]
    b = cpu_sample_size(mode, H, W)
    torch.set_num_threads(host_threads())        # torchrun exports OMP_NUM_THREADS=1; the CPU arm gets every host core
    cores = torch.get_num_threads()
    steps = max(1, min(args.steps, 3))
    fn = cpu_train_sample if mode == "train" else cpu_infer_sample
    ips, sec = fn(H, W, classes, b, steps, warmup=min(args.warmup, 1))
    sample = f"{steps} step(s) of batch {b} at {H}x{W} (of the workload's batch), torch-CPU fp32 port of the TF path, {cores} threads"
    line = {
        "impl": "reference", "metric": f"{mode} img/s", "value": round(ips, 3), "unit": "img/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": round(sec * 1e3, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "H": H, "W": W, "classes": classes, "cpu_batch": b},
        "cpu_baseline": {"value": round(ips, 3), "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(ips, 3), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "TensorFlow is not installable in this image (no wheel, no network): the CPU arm is the oracle's torch port",
    }
    if mode == "train" and not args.no_infer:      # the other half of the metric, like the B200 arm's `infer` record
        bi = cpu_sample_size("infer", H, W)
        ips_i, sec_i = cpu_infer_sample(H, W, classes, bi, steps, warmup=min(args.warmup, 1))
        line["infer"] = {"metric": "infer img/s", "value": round(ips_i, 3), "unit": "img/s", "ms_per_step": round(sec_i * 1e3, 2),
                         "e2e": {"value": round(ips_i, 3), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                         "cpu_baseline": {"value": round(ips_i, 3), "unit": "img/s", "cores": cores, "kind": "port",
                                          "sample": f"{steps} step(s) of batch {bi} at {H}x{W}, torch-CPU fp32 port, {cores} threads"}}
    _emit(line)


# ====================================================================================================== B200 arm
def synthetic_shard(batch, H, W, classes, rank):
    """synthetic shard of the global batch: uniform images, filled-rectangle masks (oracle generator's recipe restated)"""
    import torch
    g = torch.Generator(device="cuda"); g.manual_seed(2301 + rank)
    x_dev = torch.rand((batch, H, W, 3), device="cuda", generator=g)
    yy, xx = torch.meshgrid(torch.arange(H, device="cuda"), torch.arange(W, device="cuda"), indexing="ij")
    cx = (torch.rand(batch, device="cuda", generator=g) * 0.4 + 0.3) * W
    cy = (torch.rand(batch, device="cuda", generator=g) * 0.4 + 0.3) * H
    rx = (torch.rand(batch, device="cuda", generator=g) * 0.2 + 0.15) * W
    ry = (torch.rand(batch, device="cuda", generator=g) * 0.2 + 0.15) * H
    inside = ((xx[None] - cx[:, None, None]).abs() < rx[:, None, None]) & ((yy[None] - cy[:, None, None]).abs() < ry[:, None, None])
    if classes == 1:
        y_dev = inside.float()[..., None].contiguous()
    else:
        lab = (inside.long() * (1 + (xx[None] // 32 + yy[None] // 32) % (classes - 1)))
        y_dev = torch.nn.functional.one_hot(lab, classes).float().contiguous()
    return x_dev, y_dev


def roofline_of(prof, steps, hbm_peak, tc_peak, peak_kind):
    """roofline of the dominant kernel: the C-ABI entry point (one CUDA kernel, all its shapes) with the largest share of
    the step; achieved = its algorithmic bytes (or flops) over all launches / its summed launch time.  Also the per-kernel
    table and the whole-step totals."""
    rows = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])
    kernel_ms = sum(r["ms"] for _, r in rows)
    byname = {}
    for k, r in rows:
        nm = k.split("[")[0]
        if nm == "gemm_tc":
            nm = "gemm_tc(" + k.split("[")[1].split(":")[0] + ")"      # nt / convt / wgrad are different kernels
        a = byname.setdefault(nm, dict(ms=0.0, calls=0, bytes=0, flops=0, top=k, top_ms=0.0))
        a["ms"] += r["ms"]; a["calls"] += r["calls"]; a["bytes"] += r["bytes"]; a["flops"] += r["flops"]
        if r["ms"] > a["top_ms"]:
            a["top"], a["top_ms"] = k, r["ms"]
    ranked = sorted(byname.items(), key=lambda kv: -kv[1]["ms"])
    top_name, top = ranked[0]

    def family_roof(name, fam):
        ai = fam["flops"] / max(fam["bytes"], 1)
        if name.startswith("gemm_tc") and ai > RIDGE_FLOP_PER_BYTE:
            r = {"bound": "tensor", "achieved": round(fam["flops"] / (fam["ms"] * 1e-3) / 1e12, 2), "peak": tc_peak, "unit": "TFLOP/s"}
        else:
            r = {"bound": "hbm", "achieved": round(fam["bytes"] / (fam["ms"] * 1e-3) / 1e9, 1), "peak": hbm_peak, "unit": "GB/s"}
        r["frac"] = round(r["achieved"] / r["peak"], 4)
        r["kernel"] = name
        return r

    roof = family_roof(top_name, top)
    if len(ranked) > 1:     # train512: dwconv3x3_bwd and gemm_tc(nt) are within 0.2 % of the step of each other; which one leads
        ru = family_roof(*ranked[1])        # depends on the box's power-capped clock, so the second family is always shown too
        ru["share_of_step"] = round(ranked[1][1]["ms"] / kernel_ms, 4)
        roof["runner_up"] = ru
    roof["launches_per_step"] = round(top["calls"] / steps, 1)
    roof["avg_launch_ms"] = round(top["ms"] / top["calls"], 4)
    roof["algorithmic_bytes_per_launch"] = int(top["bytes"] / top["calls"])
    roof["share_of_step"] = round(top["ms"] / kernel_ms, 4)
    roof["peak_source"] = f"{peak_kind} ({'MEASURED_PEAKS.json' if peak_kind == 'measured' else 'B200_PROFILING.md fallback'}, sustained)"
    roof["traffic"] = None
    try:        # ncu --set full dram bytes of this kernel's largest shape, scaled to the average launch of the step
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f)
        shape_key = top["top"]
        if shape_key not in tr:     # another shape of the same kernel was captured
            cands = [k for k in tr if k.split("[")[0] == shape_key.split("[")[0] and k in prof]
            shape_key = cands[0] if cands else shape_key
        if shape_key in tr and shape_key in prof:
            ratio = tr[shape_key] / (prof[shape_key]["bytes"] / prof[shape_key]["calls"])
            roof["traffic"] = int(ratio * top["bytes"] / top["calls"])
            roof["traffic_over_algorithmic"] = round(ratio, 4)
            roof["traffic_shape"] = shape_key
    except Exception:
        pass
    tot_bytes = sum(r["bytes"] for _, r in rows) / steps
    tot_flops = sum(r["flops"] for _, r in rows) / steps
    table = [{"kernel": k, "calls_per_step": r["calls"] / steps, "ms_per_step": round(r["ms"] / steps, 3),
              "GBps": round(r["bytes"] / max(r["ms"], 1e-9) / 1e6, 1), "TFLOPs": round(r["flops"] / max(r["ms"], 1e-9) / 1e9, 2)}
             for k, r in rows]
    return roof, table, tot_bytes, tot_flops


def measure(model, mode, batch, H, W, classes, args, rank, local_rank, world, peaks3):
    """One workload on this rank's GPU: device-resident throughput (`value`), the live per-kernel table, and the end-to-end
    throughput through the Keras-shaped API from PAGEABLE NumPy arrays (what a user of the reference passes) and, separately,
    from caller-pinned buffers."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from unet_b200 import ops
    hbm_peak, tc_peak, peak_kind = peaks3
    eng = model.engine
    x_dev, y_dev = synthetic_shard(batch, H, W, classes, rank)

    def device_step():
        if mode == "train":      # forward + backward + (NCCL exchange) + AdamW + the compiled metrics' update (train.py:227-234)
            return model._train_step_device(x_dev, y_dev)[0]
        return eng.forward_inference(x_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()

    # ---------------- timed region 1: inputs resident in HBM (the step as shipped: CUDA-graph replay)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            device_step()
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = batch * world * args.steps / (ms_total / 1e3)

    # ---------------- the same K steps again, launched eagerly with CUDA events around every kernel launch: the live
    # per-kernel table behind `roofline` / `kernels` (event recording costs ~1-2 % of throughput, so it is kept out of `value`)
    launches0 = ops.launches
    ops.profile_begin()
    for _ in range(args.steps):
        device_step()
    prof = ops.profile_end()
    launches = ops.launches - launches0
    barrier()

    # ---------------- timed region 2: end to end through the public (Keras-shaped) API, host buffers in, host result out
    e2e = e2e_pinned = None
    if not args.no_e2e:
        x_np, y_np = x_dev.cpu().numpy(), y_dev.cpu().numpy()                   # pageable host memory, as a generator yields
        reps = max(args.steps, 2)

        def timed(fn):
            fn(2)
            barrier()
            e0.record()
            fn(args.steps)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return batch * world * args.steps / (float(t.item()) / 1e3)

        if mode == "train":
            # the call a user of the reference makes (scripts/train.py:308): model.fit(generator, steps_per_epoch=k); every
            # step stages its (x, y) NumPy batch into pinned memory, uploads it, and reads its loss back to the host
            v_page = timed(lambda k: model.fit(((x_np, y_np) for _ in range(k)), epochs=1, steps_per_epoch=k, verbose=0))
            x_pin, y_pin = torch.from_numpy(x_np).pin_memory(), torch.from_numpy(y_np).pin_memory()
            v_pin = timed(lambda k: model.fit(((x_pin, y_pin) for _ in range(k)), epochs=1, steps_per_epoch=k, verbose=0))
            h2d, d2h = x_np.nbytes + y_np.nbytes, 12
        else:
            # model.predict over k batches in ONE call (inference.py:116): per batch H2D of the images, D2H of the probabilities
            x_rep = np.empty((reps * batch, H, W, 3), np.float32)
            for r in range(reps):
                x_rep[r * batch:(r + 1) * batch] = x_np
            v_page = timed(lambda k: model.predict(x_rep[: k * batch], batch_size=batch))
            x_pin = torch.from_numpy(x_rep).pin_memory()
            v_pin = timed(lambda k: model.predict(x_pin[: k * batch], batch_size=batch))
            del x_pin
            h2d, d2h = x_np.nbytes, batch * H * W * classes * 4
        e2e = {"value": round(v_page, 2), "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "host_buffers": "pageable NumPy (staged through pinned memory inside the timed region)"}
        e2e_pinned = {"value": round(v_pin, 2), "unit": "img/s", "host_buffers": "caller-pinned"}

    rec = {"metric": f"{mode} img/s", "value": round(value, 2), "unit": "img/s", "ms_per_step": round(ms_total / args.steps, 3),
           "e2e": e2e, "e2e_pinned": e2e_pinned, "gpu_launches": launches, "clocks": clk.summary()}
    if rank == 0:
        roof, table, tot_bytes, tot_flops = roofline_of(prof, args.steps, hbm_peak, tc_peak, peak_kind)
        rec["roofline"] = roof
        rec["step_hbm"] = {"algorithmic_GB_per_step": round(tot_bytes / 1e9, 2), "GBps": round(tot_bytes / (ms_total / args.steps) / 1e6, 1),
                           "frac_of_peak": round(tot_bytes / (ms_total / args.steps) / 1e6 / hbm_peak, 4),
                           "TFLOP_per_step": round(tot_flops / 1e12, 2)}
        rec["kernels"] = table
        if args.breakdown:
            for t in table:
                print(f"{t['kernel']:58s} {t['calls_per_step']:5.1f} {t['ms_per_step']:9.3f} ms {t['GBps']:8.1f} GB/s {t['TFLOPs']:8.2f} TF/s",
                      file=sys.stderr)
    return rec


def dp_check(model, build, H, W, classes, rank, world):
    """SURVEY 8e's definition of data-parallel correctness, on NCCL: (1) the exchanged gradient == mean of the per-shard
    single-GPU gradients — checked on an fp32 replica of the model at 128x128 (exact CUDA-core arithmetic: two runs of one
    shard agree to ~1e-3, so the comparison is sharp; two bf16 runs of the same shard differ by percent-level rounding noise);
    (2) after K optimizer steps of the benchmarked model itself (bf16, CUDA-graph replay with the NCCL exchanges inside the
    graph) every rank holds bit-identical weights and BatchNormalization statistics."""
    import torch
    import torch.distributed as dist
    # ---- (1) gradient exchange == mean of shard gradients
    chk = build("train", size=128, dtype="fp32")
    e = chk.engine
    e.use_graphs, e.dropout_masks_from_step = False, False
    b = 4
    shards = [synthetic_shard(b, 128, 128, classes, 1000 + r) for r in range(world)]
    s0 = e.state.clone()
    hook, e.grad_hook = e.grad_hook, None
    acc = torch.zeros_like(e.g, dtype=torch.float64)
    for xs, ys in shards:                           # every rank computes every shard's gradient locally (no exchange) ...
        e.state.copy_(s0)
        e.train_forward_backward(xs, ys, "dice")
        acc += e.g.double()
    mean_local = acc / world
    e.grad_hook = hook
    e.state.copy_(s0)
    e.train_forward_backward(*shards[rank], "dice")    # ... then its own shard's with the exchange
    chk._grad_sync.finish()
    torch.cuda.synchronize()
    synced = e.g.double() / world
    rel = float((synced - mean_local).norm() / (mean_local.norm() + 1e-300))
    bn_avg = float((e.state - s0).abs().max())      # moving statistics rode along (averaged): they moved
    e.release_plans(); del chk, e
    # ---- (2) K real steps of the benchmarked model, then compare bit patterns across ranks
    eng = model.engine
    w0, st0 = eng.w.clone(), eng.state.clone()
    mine = synthetic_shard(b, H, W, classes, 2000 + rank)
    for _ in range(4):
        model._train_step_device(*mine)
    torch.cuda.synchronize()
    sig = torch.stack([eng.w.view(torch.int32).long().sum(), eng.state.view(torch.int32).long().sum(),
                       eng.w.double().abs().sum().view(torch.int64)])
    allsig = [torch.zeros_like(sig) for _ in range(world)]
    dist.all_gather(allsig, sig)
    same = all(bool((a == allsig[0]).all()) for a in allsig)
    moved = bool((eng.w != w0).any())
    eng.w.copy_(w0); eng.state.copy_(st0); eng._stage_dirty = True
    eng.reset_optimizer()
    eng.release_plans()
    torch.cuda.empty_cache()
    return {"grad_equals_mean_of_shard_grads_rel_err": rel, "grad_ok": rel < 5e-3, "grad_check": "fp32 replica, 128x128, batch 4 per rank (two runs of one shard agree to ~1e-3: fp32 atomics order through 18 BatchNormalization layers; a missing or wrong exchange is off by > 0.3)",
            "bn_statistics_exchanged": bn_avg > 0.0,
            "weights_identical_across_ranks_after_4_steps": same, "weights_moved": moved, "world": world, "shard_batch": b}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from unet_b200 import dist as D
    from unet_b200.keras_api import AdamW, MeanIoU, Model
    from utils.metrics import dice_coef

    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line
    rank, local_rank, world = D.init_from_env("nccl")
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    torch.cuda.set_device(local_rank)
    if world > 1:       # torchrun pins OMP_NUM_THREADS=1: give each rank its share of the host cores for the staging copies
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    mode, H, W, classes, batch, desc = WORKLOADS[args.workload]
    batch = args.batch or batch
    P3 = peaks()

    def build(mode_, size=None, dtype=None):
        m = Model((size or H, size or W, 3), num_classes=classes, dropout_rate=0.2, use_batch_norm=True, dtype=dtype or args.dtype, seed=2301)
        # the reference's compile call (scripts/train.py:226-234): AdamW + dice_loss + [MeanIoU(2, 'mean_io_u'), dice_coef]
        m.compile(optimizer=AdamW(learning_rate=2e-3, weight_decay=1e-4), loss="dice_loss",
                  metrics=[MeanIoU(num_classes=max(2, classes), name="mean_io_u"), dice_coef])
        if args.no_graphs:
            m.engine.use_graphs = False
        if world > 1:
            dist.broadcast(m.engine.w, src=0)
            dist.broadcast(m.engine.state, src=0)
            m.engine._stage_dirty = True
            if mode_ == "train":
                m.enable_data_parallel()
        return m

    model = build(mode)
    check = None
    if args.check and world > 1 and mode == "train":
        check = dp_check(model, build, H, W, classes, rank, world)
    rec = measure(model, mode, batch, H, W, classes, args, rank, local_rank, world, P3)
    eng = model.engine
    graph = bool(eng.use_graphs and (model._grad_sync is None or getattr(eng, "graph_collectives", False)))

    infer = None
    if mode == "train" and not args.no_infer:
        # the other half of BASELINE.json's metric ("train & infer img/s"): inference at the same resolution and batch, on
        # the same model object (weights as trained so far), no communication (each rank serves its shard)
        eng.release_plans()
        torch.cuda.empty_cache()
        hook, eng.grad_hook = eng.grad_hook, None
        infer = measure(model, "infer", batch, H, W, classes, args, rank, local_rank, world, P3)
        eng.grad_hook = hook
        infer["config"] = {"workload": f"U-Net {H}x{W} inference, batch {batch}/GPU (model.predict, scripts/inference.py:116)",
                           "batch_per_gpu": batch, "parallelism": f"batch sharded over {world} GPU(s), no communication"}
        if rank == 0:
            infer["kernels"] = infer["kernels"][:6]

    # CUDA graphs that captured NCCL kernels must be gone before the process group is destroyed (destroy_process_group
    # otherwise waits on them forever)
    eng.release_plans()
    torch.cuda.synchronize()
    if rank != 0:
        return
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(host_threads())
        b = cpu_sample_size(mode, H, W)
        fn = cpu_train_sample if mode == "train" else cpu_infer_sample
        ips, sec = fn(H, W, classes, b, 2, warmup=1)
        cpu = {"value": round(ips, 3), "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"2 steps of batch {b} at {H}x{W}, torch-CPU fp32 port of the reference's TF path ({sec:.1f} s/step)"}

    line = {
        "metric": rec["metric"], "value": rec["value"], "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": desc, "H": H, "W": W, "classes": classes, "batch_per_gpu": batch,
                   "global_batch": batch * world, "dropout": 0.2, "optimizer": "AdamW(2e-3, wd 1e-4)" if mode == "train" else None,
                   "metrics": "[MeanIoU(2,'mean_io_u'), dice_coef] updated every step" if mode == "train" else None,
                   "parallelism": f"dp{world}", "cuda_graph": graph,
                   "l2": "inputs and activations exceed L2 (126 MB) many times over; no flush needed"},
        "clocks": rec["clocks"], "e2e": rec["e2e"], "e2e_pinned": rec["e2e_pinned"], "gpu_launches": rec["gpu_launches"],
        "roofline": rec["roofline"], "cpu_baseline": cpu, "step_hbm": rec["step_hbm"], "kernels": rec["kernels"][:12],
    }
    if check is not None:
        line["check"] = check
    if infer is not None:
        line["infer"] = infer          # last key: the end of the line is what a log tail shows
    _emit(line)


_JSON_FD = None


def _reserve_stdout():
    """stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner to fd 1) are sent to stderr for
    the whole run and the line is written to the original stdout at the end."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    args = parse()
    if os.environ.get("BENCH_FAULT_TIMEOUT"):      # debugging aid: dump every thread's Python stack if the run is still going
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["BENCH_FAULT_TIMEOUT"]), exit=True, file=sys.stderr)
    _reserve_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    run_b200(args)
    try:
        from unet_b200 import dist as D
        # the JSON line is out; never let communicator teardown hold the process (and the driver's timer) hostage
        if not D.shutdown(timeout_s=20.0):
            sys.stderr.write("bench.py: destroy_process_group did not return within 20 s; exiting\n")
            sys.stderr.flush()
            os._exit(0)
    except Exception:
        pass


if __name__ == "__main__":
    main()
