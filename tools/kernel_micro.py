#!/usr/bin/env python3
"""Run single kernels of libunet_b200 at the level-0 shapes of the train512 workload (batch 64, 512x512) — the
subject for `ncu --set full` captures and quick CUDA-event timings.  usage: kernel_micro.py <name> [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200 import ops

name = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
B, H, W = 64, 512, 512
M = B * H * W
bf = torch.bfloat16
dev = "cuda"


def rnd(*shape, dtype=bf):
    return (torch.rand(shape, device=dev) - 0.5).to(dtype)


def run(fn, nbytes):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms:.3f} ms  {nbytes / ms / 1e6:.0f} GB/s")


if name == "gemm64":          # pointwise 64->64 forward with BN statistics (enc1_block2 / dec1_block2)
    A, Bt, C = rnd(M, 64), rnd(64, 64), torch.empty((M, 64), device=dev, dtype=bf)
    cs = torch.zeros(64, device=dev, dtype=torch.float64); cq = torch.zeros_like(cs)
    run(lambda: ops.gemm(A, Bt, C, b_trans=True, epilogue=ops.EPI_STATS, colsum=cs, colsq=cq), 2 * M * 64 * 2)
elif name == "gemm64_none":
    A, Bt, C = rnd(M, 64), rnd(64, 64), torch.empty((M, 64), device=dev, dtype=bf)
    run(lambda: ops.gemm(A, Bt, C, b_trans=True), 2 * M * 64 * 2)
elif name == "gemm128_64":    # dec1_block1 pointwise 128->64
    A, Bt, C = rnd(M, 128), rnd(64, 128), torch.empty((M, 64), device=dev, dtype=bf)
    run(lambda: ops.gemm(A, Bt, C, b_trans=True), M * 192 * 2)
elif name == "convt":         # dec1_upsample
    x, Bt = rnd(B, 256, 256, 128), rnd(256, 128)
    cat = torch.empty((B, 512, 512, 128), device=dev, dtype=bf)
    bias = torch.zeros(64, device=dev)
    run(lambda: ops.gemm(x, Bt, cat[..., :64], b_trans=True, epilogue=ops.EPI_CONVT, shift=bias, convt_hw=(256, 256)),
        (x.numel() + M * 64) * 2)
elif name == "gemm_convt_shape":   # the dec1_upsample contraction with a plain row-major store (isolates the 5-D pixel-shuffle store)
    x, Bt = rnd(B * 256 * 256, 128), rnd(256, 128)
    C = torch.empty((B * 256 * 256, 256), device=dev, dtype=bf)
    bias = torch.zeros(256, device=dev)
    run(lambda: ops.gemm(x, Bt, C, b_trans=True, epilogue=ops.EPI_AFFINE, shift=bias), (x.numel() + C.numel()) * 2)
elif name in ("pw_bneck2", "pw_dec4b1", "pw_enc4b2", "dgrad_bneck2"):
    # tensor-bound pointwise GEMMs of train512: bneck_block2.pw 1024->1024 at 32x32, dec4_block1.pw 1024->512 and
    # enc4_block2.pw 512->512 at 64x64 (forward with BatchNormalization statistics), and bneck_block2's data gradient
    m, k, n = {"pw_bneck2": (B * 32 * 32, 1024, 1024), "pw_dec4b1": (B * 64 * 64, 1024, 512), "pw_enc4b2": (B * 64 * 64, 512, 512),
               "dgrad_bneck2": (B * 32 * 32, 1024, 1024)}[name]
    A, Bt, C = rnd(m, k), rnd(n, k), torch.empty((m, n), device=dev, dtype=bf)
    cs = torch.zeros(n, device=dev, dtype=torch.float64); cq = torch.zeros_like(cs)
    flops = 2 * m * k * n
    if name.startswith("dgrad"):
        fn = lambda: ops.gemm(A, Bt, C, b_trans=True)
    else:
        fn = lambda: ops.gemm(A, Bt, C, b_trans=True, epilogue=ops.EPI_STATS, colsum=cs, colsq=cq)
    run(fn, (m * k + m * n + n * k) * 2)
    print(f"{name}: {flops / 1e12:.3f} TFLOP per launch")
elif name in ("convt_dec4", "convt_dec3", "convt_dec2", "convt_dec4_nodrop", "convt_dec3_nodrop", "convt_dec2_nodrop"):
    # Conv2DTranspose GEMMs with the 5-D pixel-shuffle store into the concat buffer (+ Dropout in the epilogue, as trained)
    hh, cin = {"convt_dec4": (32, 1024), "convt_dec3": (64, 512), "convt_dec2": (128, 256)}[name.replace("_nodrop", "")]
    cout = cin // 2
    x, Bt = rnd(B, hh, hh, cin), rnd(4 * cout, cin)
    cat = torch.empty((B, 2 * hh, 2 * hh, 2 * cout), device=dev, dtype=bf)
    bias = torch.zeros(cout, device=dev)
    drop = None if name.endswith("_nodrop") else ops.make_dropout(0.2, 5, ctot=2 * cout, c0=0)
    run(lambda: ops.gemm(x, Bt, cat[..., :cout], b_trans=True, epilogue=ops.EPI_CONVT, shift=bias, convt_hw=(hh, hh), drop=drop),
        (x.numel() + B * 4 * hh * hh * cout + Bt.numel()) * 2)
    print(f"{name}: {2 * B * hh * hh * cin * 4 * cout / 1e12:.3f} TFLOP per launch")
elif name in ("gemm64_fp32", "gemm512_fp32", "wgrad64_fp32"):   # fp32 mode on the tensor cores (tf32x3)
    f32 = torch.float32
    if name == "wgrad64_fp32":
        A, Bm, C = rnd(M // 2, 64, dtype=f32), rnd(M // 2, 64, dtype=f32), torch.zeros((64, 64), device=dev)
        run(lambda: ops.gemm(A, Bm, C, a_trans=True, accumulate=True, tf32x3=True, tensor_core=True), 2 * (M // 2) * 64 * 4 * 2)
    else:
        m, k, n = (M // 2, 64, 64) if name == "gemm64_fp32" else (B * 64 * 64, 512, 512)
        A, W, C = rnd(m, k, dtype=f32), rnd(n, k, dtype=f32), torch.empty((m, n), device=dev)
        Wl = torch.empty_like(W); ops.split_tf32(W, None, Wl)
        Al = ops.tf32_lo(A, slot=3)
        run(lambda: ops.gemm(A, W, C, b_trans=True, B_lo=Wl, A_lo=Al, tensor_core=True), (2 * m * k + m * n) * 4)
        print(f"{name}: {2 * m * k * n / 1e12:.3f} TFLOP per launch (fp32-equivalent)")
elif name == "wgrad64":
    A, Bm, C = rnd(M, 64), rnd(M, 64), torch.zeros((64, 64), device=dev)
    run(lambda: ops.gemm(A, Bm, C, a_trans=True, accumulate=True), 2 * M * 64 * 2)
elif name == "dw_fwd":
    x, y, w = rnd(B, H, W, 64), torch.empty((B, H, W, 64), device=dev, dtype=bf), torch.rand((9, 64), device=dev)
    run(lambda: ops.dwconv3x3(x, w, y), 2 * M * 64 * 2)
elif name == "dw_fwd128":
    x, y, w = rnd(B, H, W, 128), torch.empty((B, H, W, 128), device=dev, dtype=bf), torch.rand((9, 128), device=dev)
    run(lambda: ops.dwconv3x3(x, w, y), 2 * M * 128 * 2)
elif name == "dw_fwd256":
    x, y, w = rnd(B, 256, 256, 256), torch.empty((B, 256, 256, 256), device=dev, dtype=bf), torch.rand((9, 256), device=dev)
    run(lambda: ops.dwconv3x3(x, w, y), 2 * B * 256 * 256 * 256 * 2)
elif name in ("dw_fwd_aff", "dw_fwd_aff128_256", "dw_fwd_aff256_128"):   # *_block2 depthwise: BN+ReLU of block1 applied on load, colsum
    cc, hh = {"dw_fwd_aff": (64, 512), "dw_fwd_aff128_256": (128, 256), "dw_fwd_aff256_128": (256, 128)}[name]
    x, y, w = rnd(B, hh, hh, cc), torch.empty((B, hh, hh, cc), device=dev, dtype=bf), torch.rand((9, cc), device=dev)
    sc, sh, cs = torch.rand(cc, device=dev) + 0.5, torch.rand(cc, device=dev) - 0.5, torch.zeros(cc, device=dev)
    run(lambda: ops.dwconv3x3(x, w, y, in_scale=sc, in_shift=sh, colsum=cs), 2 * B * hh * hh * cc * 2)
elif name == "dw_bwd_w":
    x, dy, dw = rnd(B, H, W, 64), rnd(B, H, W, 64), torch.zeros((9, 64), device=dev)
    run(lambda: ops.dwconv3x3_bwd_weight(x, dy, dw), 2 * M * 64 * 2)
elif name == "gemm_fold_dgrad":   # dec1_block1 data gradient with BatchNormalization backward folded in: dd = [g | z] * wab^T + bias
    g_, z_, wab = rnd(M, 64), rnd(M, 64), rnd(128, 128)
    bias, dd = torch.rand(128, device=dev), torch.empty((M, 128), device=dev, dtype=bf)
    run(lambda: ops.gemm(g_, wab, dd, b_trans=True, A2=z_, epilogue=ops.EPI_AFFINE, shift=bias), M * (64 + 64 + 128) * 2)
elif name == "gemm_fold_wgrad":   # its weight gradient: G = d^T [g | z]
    d_, g_, z_ = rnd(M, 128), rnd(M, 64), rnd(M, 64)
    G = torch.zeros((128, 128), device=dev)
    run(lambda: ops.gemm(d_, g_, G, a_trans=True, accumulate=True, B2=z_), M * (128 + 64 + 64) * 2)
elif name in ("pw_bwd_fused64", "pw_bwd_fused128"):   # folded pointwise backward, both contractions from one pass (level 0)
    cin = 64 if name.endswith("64") else 128
    g_, z_, d_ = rnd(M, 64), rnd(M, 64), rnd(M, cin)
    wab, bias = rnd(cin, 128), torch.rand(cin, device=dev)
    dd, G = torch.empty((M, cin), device=dev, dtype=bf), torch.zeros((cin, 128), device=dev)
    run(lambda: ops.pw_bwd_fused(g_, z_, d_, wab, bias, dd, G), M * (128 + 2 * cin) * 2)
elif name in ("dw_bwd_aff", "dw_bwd128", "dw_bwd_drop256", "dw_bwd_aff128_256", "dw_bwd128_up", "dw_bwd_drop256_up"):
    cc, hh = {"dw_bwd_aff": (64, 512), "dw_bwd128": (128, 512), "dw_bwd_drop256": (256, 256), "dw_bwd_aff128_256": (128, 256),
              "dw_bwd128_up": (128, 512), "dw_bwd_drop256_up": (256, 256)}[name]
    x, dy, dw = rnd(B, hh, hh, cc), rnd(B, hh, hh, cc), torch.zeros((9, cc), device=dev)
    dx, w, sums = torch.empty_like(x), torch.rand((9, cc), device=dev), torch.zeros((2, cc), device=dev)
    sc, sh = torch.rand(cc, device=dev) + 0.5, torch.rand(cc, device=dev) - 0.5
    nb = 3 * B * hh * hh * cc * 2
    if name.endswith("_up"):     # dec*_block1: upsampled half of the concat gradient stored un-pixel-shuffled + bias gradient
        gth, db = torch.empty((B * hh * hh // 4, 2 * cc), device=dev, dtype=bf), torch.zeros(cc // 2, device=dev)
        drop = ops.make_dropout(0.2, 5, ctot=cc, c0=0) if "drop" in name else None
        run(lambda: ops.dwconv3x3_bwd(x, dy, w, dx, dw, drop=drop, up_out=gth, up_colsum=db), nb)
    elif "aff" in name:
        run(lambda: ops.dwconv3x3_bwd(x, dy, w, dx, dw, relu_mask=True, bn_sums=sums, x_scale=sc, x_shift=sh), nb)
    elif "drop" in name:
        drop = ops.make_dropout(0.2, 5, ctot=cc, c0=0)
        run(lambda: ops.dwconv3x3_bwd(x, dy, w, dx, dw, drop=drop, drop_c_from=cc // 2), nb)
    else:
        run(lambda: ops.dwconv3x3_bwd(x, dy, w, dx, dw), nb)
elif name in ("dw_bwd", "dw_bwd_mask"):
    x, dy, dw = rnd(B, H, W, 64), rnd(B, H, W, 64), torch.zeros((9, 64), device=dev)
    dx, w, sums = torch.empty_like(x), torch.rand((9, 64), device=dev), torch.zeros((2, 64), device=dev)
    mask = name.endswith("mask")
    run(lambda: ops.dwconv3x3_bwd(x, dy, w, dx, dw, relu_mask=mask, bn_sums=sums if mask else None), 3 * M * 64 * 2)
elif name in ("bn_act_pool", "bn_act_pool_cat", "bn_act_cat", "maxpool_bwd"):
    # encoder block2 activation: y (optionally into the skip half of the concat buffer, ld = 2C) + the 2x2 max-pooled tensor
    z = rnd(B, H, W, 64)
    sc, sh = torch.rand(64, device=dev), torch.rand(64, device=dev)
    cat = torch.empty((B, H, W, 128), device=dev, dtype=bf)
    y = cat[..., 64:] if name.endswith("cat") else torch.empty((B, H, W, 64), device=dev, dtype=bf)
    pooled = torch.empty((B, H // 2, W // 2, 64), device=dev, dtype=bf)
    if name == "maxpool_bwd":
        dpool, dy, sums = rnd(B, H // 2, W // 2, 64), torch.empty((B, H, W, 64), device=dev, dtype=bf), torch.zeros((2, 64), device=dev)
        dcat = rnd(B, H, W, 128)
        run(lambda: ops.maxpool2x2_bwd(z, sc, sh, dpool, dcat[..., 64:], dy, bn_sums=sums), int(3.25 * M * 64 * 2))
    elif name == "bn_act_cat":
        run(lambda: ops.bn_act(z, sc, sh, y, relu=True), 2 * M * 64 * 2)
    else:
        run(lambda: ops.bn_act(z, sc, sh, y, relu=True, pooled=pooled), int(2.25 * M * 64 * 2))
elif name in ("bn_bwd_reduce", "bn_bwd_apply", "bn_act"):
    z, dy, dz = rnd(B, H, W, 64), rnd(B, H, W, 64), torch.empty((B, H, W, 64), device=dev, dtype=bf)
    v = lambda: torch.rand(64, device=dev)
    sc, sh, mu, rs, dg, db = v(), v(), v(), v(), torch.zeros(64, device=dev), torch.zeros(64, device=dev)
    if name == "bn_bwd_reduce":
        run(lambda: ops.bn_bwd_reduce(dy, z, sc, sh, mu, rs, dg, db), 2 * M * 64 * 2)
    elif name == "bn_bwd_apply":
        run(lambda: ops.bn_bwd_apply(dy, z, sc, sh, mu, rs, dg, db, dz), 3 * M * 64 * 2)
    else:
        run(lambda: ops.bn_act(z, sc, sh, dz), 2 * M * 64 * 2)
elif name in ("fused64", "fused64_ns", "fused64_pool", "fused64_head", "fused128_64", "fused64_128", "fused256_128", "fused128_128"):
    # inference conv_block kernel; _ns: BatchNormalization scale folded into the pointwise kernel (what the engine runs)
    cin, cout, hh = {"fused64": (64, 64, 512), "fused64_ns": (64, 64, 512), "fused64_pool": (64, 64, 512), "fused64_head": (64, 64, 512), "fused128_64": (128, 64, 512),
                     "fused64_128": (64, 128, 256), "fused256_128": (256, 128, 256), "fused128_128": (128, 128, 256)}[name]
    x = rnd(B, hh, hh, cin); y = torch.empty((B, hh, hh, cout), device=dev, dtype=bf)
    wd = torch.rand((9, cin), device=dev); wpt = rnd(cout, cin)
    sc, sh = torch.rand(cout, device=dev), torch.rand(cout, device=dev)
    if name == "fused64":
        run(lambda: ops.sepconv_fused(x, wd, wpt, y, scale=sc, shift=sh), (x.numel() + y.numel()) * 2)
    elif name == "fused64_head":            # dec1_block2 + output head: only the probabilities leave the chip
        hw, hb = torch.rand((cout, 1), device=dev) - 0.5, torch.zeros(1, device=dev)
        probs = torch.empty((B, hh, hh, 1), device=dev)
        run(lambda: ops.sepconv_fused(x, wd, wpt, None, shift=sh, head_w=hw, head_b=hb, head_out=probs), x.numel() * 2 + probs.numel() * 4)
    elif name == "fused64_pool":
        pooled = torch.empty((B, hh // 2, hh // 2, cout), device=dev, dtype=bf)
        run(lambda: ops.sepconv_fused(x, wd, wpt, y, shift=sh, pooled=pooled), (x.numel() + y.numel() + pooled.numel()) * 2)
    else:
        run(lambda: ops.sepconv_fused(x, wd, wpt, y, shift=sh), (x.numel() + y.numel()) * 2)
elif name in ("head8_fwd", "head8_bwd", "head1_fwd", "head1_bwd"):
    C = 8 if "8" in name else 1
    Bh = 32
    Mh = Bh * H * W
    x = rnd(Bh, H, W, 64)
    wk, bk = torch.rand((64, C), device=dev) - 0.5, torch.zeros(C, device=dev)
    probs = torch.empty((Bh, H, W, C), device=dev)
    yt = (torch.rand((Bh, H, W, C), device=dev) > 0.5).float()
    sums = torch.zeros((Bh, C, 3), device=dev, dtype=torch.float64)
    coef = torch.rand((Bh, C, 2), device=dev)
    dx = torch.empty_like(x); dw = torch.zeros((64, C), device=dev); db = torch.zeros(C, device=dev)
    ops.head_fwd(x, wk, bk, probs, yt, sums)
    if name.endswith("fwd"):
        run(lambda: ops.head_fwd(x, wk, bk, probs, yt, sums), Mh * (128 + 8 * C))
    else:
        run(lambda: ops.head_bwd(x, wk, probs, yt, coef, dx, dw, db), Mh * (256 + 8 * C))
else:
    raise SystemExit(f"unknown kernel {name}")
