#!/usr/bin/env python3
"""Same-process A/B of engine schedule switches on the train512 workload (batch 64, 512x512, bf16, CUDA-graph replay):
clocks under the power cap differ between boxes and runs, so variants are only comparable inside one process.
usage: ab_step.py [steps] — prints ms/step per variant, interleaved over 3 rounds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200.engine import UNetEngine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
B, H, W = 64, 512, 512
VARIANTS = {
    "all on": {},
    "direct convT bwd with dropout": {"convt_bwd_direct_drop": True},
    "no fuse_pw_bwd": {"fuse_pw_bwd": False},
    "no convt_bwd_direct": {"convt_bwd_direct": False},
    "no fold_bn_bwd": {"fold_bn_bwd": False},
    "no fuse_bn_act": {"fuse_bn_act": False},
    "no defer_dropout": {"defer_dropout": False},
    "no fuse_dw_bwd": {"fuse_dw_bwd": False, "fold_bn_bwd": False},
}
g = torch.Generator(device="cuda").manual_seed(2301)
x = torch.rand((B, H, W, 3), device="cuda", generator=g)
y = (torch.rand((B, H, W, 1), device="cuda", generator=g) > 0.7).float()
eng = UNetEngine((H, W, 3), dtype="bf16")
eng.use_graphs = True
res = {k: [] for k in VARIANTS}
for rnd in range(3):
    for name, flags in VARIANTS.items():
        for k in ("fold_bn_bwd", "defer_dropout", "fuse_dw_bwd", "fuse_bn_act", "fuse_pw_bwd", "convt_bwd_direct"):
            setattr(eng, k, flags.get(k, True))
        eng.convt_bwd_direct_drop = flags.get("convt_bwd_direct_drop", False)
        eng.release_plans()
        for _ in range(3):
            eng.train_step(x, y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            eng.train_step(x, y)
        e1.record(); torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / steps)
for name, v in res.items():
    print(f"{name:30s} " + "  ".join(f"{t:7.3f}" for t in v) + f"   min {min(v):7.3f} ms  -> {B / min(v) * 1e3:7.1f} img/s")
