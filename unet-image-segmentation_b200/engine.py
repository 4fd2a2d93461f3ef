"""The B200 execution engine for the reference U-Net (model/u_net.py:28-116): parameters, buffer plan and the
forward / backward / optimizer schedules, expressed as sequences of C-ABI kernel launches (ops.py).

Replaces, for this one model, what Keras + TensorFlow do behind `model.predict` / `model.fit`
(scripts/inference.py:116, scripts/train.py:308): layer graph execution, autodiff, AdamW (train.py:226).

Layout in HBM
  * activations NHWC, bf16 (tcgen05 path) or fp32 (parity path); probabilities, parameters, statistics fp32.
  * Concatenate (u_net.py:96) is zero-copy: Conv2DTranspose writes channels [0,f) and the encoder skip writes
    channels [f,2f) of one [N,h,w,2f] buffer.
  * trainable parameters live in ONE flat fp32 buffer in Keras creation order (grad / Adam m / Adam v mirror it), so
    AdamW is one launch and the data-parallel gradient exchange is a few contiguous all-reduces.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from .spec import BN_EPS, BN_MOMENTUM, FILTERS, UNetSpec

SMOOTH = 1e-7   # K.epsilon(), utils/metrics.py:4


class _Plan:
    """Named device buffers for one (batch, mode) configuration; allocated once, reused every step (graph-safe)."""

    def __init__(self, device, act_dtype):
        self.device, self.act_dtype = device, act_dtype
        self.t: Dict[str, torch.Tensor] = {}

    def buf(self, name: str, shape, dtype=None, zero=False) -> torch.Tensor:
        t = self.t.get(name)
        if t is None:
            dtype = dtype or self.act_dtype
            t = (torch.zeros if zero else torch.empty)(tuple(shape), device=self.device, dtype=dtype)
            self.t[name] = t
        return t

    def bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.t.values())


class UNetEngine:
    def __init__(self, input_size, num_classes: int = 1, dropout_rate: float = 0.2, use_batch_norm: bool = True,
                 dtype: str = "bf16", device: Optional[torch.device] = None, seed: int = 2301):
        self.spec = UNetSpec(tuple(input_size), num_classes, dropout_rate, use_batch_norm)
        H, W, _ = self.spec.input_size
        if H % 16 or W % 16:
            raise ValueError(f"input height and width must be multiples of 16 (4 MaxPooling2D stages), got {H}x{W}")
        if num_classes < 1 or num_classes > 8:
            raise ValueError("num_classes must be in 1..8")
        if dtype not in ("bf16", "fp32"):
            raise ValueError("dtype must be 'bf16' or 'fp32'")
        if not torch.cuda.is_available():
            raise RuntimeError("unet_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        ops.device_check(self.device.index or 0)
        self.dtype_name = dtype
        self.act_dtype = torch.bfloat16 if dtype == "bf16" else torch.float32
        self.num_classes, self.dropout_rate, self.use_bn = num_classes, float(dropout_rate), bool(use_batch_norm)
        sp = self.spec
        dev = self.device
        with torch.cuda.device(dev):
            self.w = torch.zeros(sp.n_trainable_flat, device=dev)
            self.g = torch.zeros_like(self.w)
            self.m = torch.zeros_like(self.w)
            self.v = torch.zeros_like(self.w)
            self.state = torch.zeros(max(sp.n_state_flat, 8), device=dev)
            # per-BN vectors: scale, shift (batch or folded), saved mean, saved rstd; fp64 column sums for statistics
            self._bn_off: Dict[str, Tuple[int, int]] = {}
            tot = 0
            for b in sp.blocks:
                self._bn_off[b.prefix] = (tot, b.cout)
                tot += b.cout
            self._bn_total = tot
            self.bn_vec = torch.zeros((4, tot), device=dev)
            self.fold = torch.zeros((2, tot), device=dev)
            self.colstats = torch.zeros((2, tot), device=dev, dtype=torch.float64)
            self.ones = torch.ones(max(FILTERS) * 2, device=dev)
            self.zeros = torch.zeros(max(FILTERS) * 2, device=dev)
            # lr, wd, beta1, beta2, eps, t, grad_scale, unused
            self.hyper = torch.tensor([2e-3, 1e-4, 0.9, 0.999, 1e-7, 1.0, 1.0, 0.0], device=dev)
            self.step_word = torch.zeros(1, device=dev, dtype=torch.int32)
            self.fold_guard = torch.zeros(1, device=dev, dtype=torch.int32)    # set by bn_bwd_coef: the fold is ill-conditioned
            # folded BN backward (per block): sums [2,Cout] | sd [Cin] | G [Cin,2Cout] live in ONE fp32 buffer zeroed per step
            self._fz_off: Dict[str, Tuple[int, int, int]] = {}
            tot = 0
            for b in sp.blocks:
                if b.prefix.endswith("_block1") or b.prefix.startswith("enc") or b.prefix == "dec1_block2":
                    self._fz_off[b.prefix] = (tot, b.cin, b.cout)
                    tot += 2 * b.cout + b.cin + 2 * b.cin * b.cout
                    tot = (tot + 63) // 64 * 64
            self._fz = torch.zeros(max(tot, 64), device=dev)
            self._fold_w: Dict[str, Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = {}
        self._stage: Dict[str, torch.Tensor] = {}
        self._stage_table = None        # (device table, n, tiles) of ops.cast_transpose_bf16_batched over self.w
        self._stage32: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}    # fp32 mode: tf32 (hi, lo) parts of the dense kernels
        import os as _os
        self.fp32_tensor_cores = _os.environ.get("UNET_B200_FP32_TC", "1") != "0"   # fp32 mode: contractions as 3 x tf32 on tcgen05
        self._stage_fold: Dict[str, torch.Tensor] = {}     # inference: bf16 [Cout, Cin] pointwise kernels with the BN scale folded in
        self.fold_scale_into_weights = True
        self._stage_dirty = True
        self._fold_dirty = True
        self._plans: Dict[Tuple[int, bool], _Plan] = {}
        self._drop_seed = {name: (seed * 7919 + i * 104729) & 0x7FFFFFFF
                           for i, name in enumerate(["bneck_dropout", "dec4_dropout", "dec3_dropout", "dec2_dropout"])}
        self.dropout_masks_from_step = True     # False: masks depend only on the seeds (parity tests)
        self.fuse_sepconv = True                # inference: levels with <= 128 output channels run the fused conv_block kernel
        self.fuse_dw_bwd = True                 # training: depthwise input + weight gradients from one pass over dy
        self.fuse_pool = True                   # inference: MaxPooling2D from the fused conv_block kernel's staged tile
        self.convt_bwd_direct = True            # training: no un-pixel-shuffle gather pass (the depthwise backward stores that layout)
        self.convt_bwd_direct_drop = False      # ... also where the concat carries Dropout (dec2-4): the depthwise kernel then masks both halves
        self.fuse_pw_bwd = True                 # training: folded data + weight gradient of a 64-channel pointwise from one pass
        self.fold_bn_bwd = True                 # training (bf16, BN): BatchNormalization backward folded into the block's pointwise
                                                # data / weight gradient GEMMs (no reduce / apply passes, no dz tensor) wherever the
                                                # producer of dy has the block's activation in registers: every *_block1 (depthwise
                                                # backward of *_block2), enc*_block2 (max-pool backward), dec1_block2 (head backward)
        self.fuse_bn_act = True                 # training: the BN+ReLU pass of every *_block1 is not run; *_block2's depthwise kernels
                                                # (forward and fused backward) read the pre-BN tensor and apply it on load
        self.defer_dropout = True               # training: the concat-buffer Dropout mask of dcat is applied by its readers
        self.fuse_head = True                   # inference: output head fused into dec1_block2's GEMM epilogue (bf16 path)
        self.use_graphs = False                 # replay inference / single-GPU training steps from CUDA graphs
        self._graphs: Dict[tuple, tuple] = {}
        self.grad_hook = None                   # callable(region) — dist.GradSync.ready; regions: decoder, bottleneck, encoder
        self.grad_finish = None                 # callable() — dist.GradSync.finish: the compute stream joins the exchanges
        self.graph_collectives = True           # data parallel: capture the step INCLUDING its NCCL exchanges into the CUDA graph
        self.init_weights(seed)

    # ------------------------------------------------------------------------------------------------ parameters
    def wview(self, name: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        p = self.spec.params[name]
        base = (self.w if buf is None else buf) if p.trainable else self.state
        return base[p.offset:p.offset + p.size]

    def _mat(self, name: str, buf: Optional[torch.Tensor] = None) -> torch.Tensor:
        """2-D GEMM view of a kernel in its Keras memory order."""
        p = self.spec.params[name]
        v = self.wview(name, buf)
        leaf = name.split("/")[1]
        if leaf == "depthwise_kernel":
            return v.view(9, p.shape[2])
        if leaf == "pointwise_kernel":
            return v.view(p.shape[2], p.shape[3])                 # [Cin, Cout]
        if leaf == "kernel" and len(p.shape) == 4 and p.shape[0] == 2:
            return v.view(4 * p.shape[2], p.shape[3])             # convT: [(a,b,co), Cin]
        if leaf == "kernel":
            return v.view(p.shape[2], p.shape[3])                 # head: [64, classes]
        return v

    def init_weights(self, seed: int = 2301) -> None:
        """Keras default initialisers (Glorot-uniform kernels, zero biases, gamma 1, beta 0, moving 0/1)."""
        rng = np.random.default_rng(seed)
        out = {}
        for name, p in self.spec.params.items():
            leaf = name.split("/")[1]
            if leaf.endswith("kernel"):
                rf = int(np.prod(p.shape[:-2]))
                lim = np.sqrt(6.0 / (rf * p.shape[-2] + rf * p.shape[-1]))
                out[name] = rng.uniform(-lim, lim, size=p.shape).astype(np.float32)
            elif leaf in ("gamma", "moving_variance"):
                out[name] = np.ones(p.shape, np.float32)
            else:
                out[name] = np.zeros(p.shape, np.float32)
        self.set_weights(out)

    def set_weights(self, weights: Dict[str, np.ndarray]) -> None:
        """name -> array in Keras shape.  Unknown names raise; missing names keep their value."""
        for name, arr in weights.items():
            if name not in self.spec.params:
                raise KeyError(f"unknown weight {name!r}")
            p = self.spec.params[name]
            a = np.ascontiguousarray(np.asarray(arr, dtype=np.float32))
            if tuple(a.shape) != tuple(p.shape):
                raise ValueError(f"{name}: expected shape {p.shape}, got {a.shape}")
            self.wview(name).copy_(torch.from_numpy(a.reshape(-1)))
        self._stage_dirty = True
        self._fold_dirty = True

    def get_weights(self) -> Dict[str, np.ndarray]:
        return {name: self.wview(name).detach().cpu().numpy().reshape(p.shape).copy()
                for name, p in self.spec.params.items()}

    def reset_optimizer(self) -> None:
        self.m.zero_(); self.v.zero_()
        self.hyper[5] = 1.0

    def set_hyper(self, lr=None, weight_decay=None, beta1=None, beta2=None, eps=None, grad_scale=None) -> None:
        for i, val in enumerate((lr, weight_decay, beta1, beta2, eps)):
            if val is not None:
                self.hyper[i] = float(val)
        if grad_scale is not None:
            self.hyper[6] = float(grad_scale)

    # bf16 operand staging for the tensor-core path
    def _restage(self) -> None:
        if not self._stage_dirty:
            return
        if self.act_dtype == torch.bfloat16:
            if self._stage_table is None:     # every dense kernel in one launch: the table of (source offset, destinations) is fixed
                items = []
                for name, p in self.spec.params.items():
                    leaf = name.split("/")[1]
                    if leaf == "pointwise_kernel" or (leaf == "kernel" and p.shape[0] == 2):
                        src = self._mat(name)
                        r, c = src.shape
                        a = self._stage[name] = torch.empty((r, c), device=self.device, dtype=torch.bfloat16)
                        at = self._stage[name + "^T"] = torch.empty((c, r), device=self.device, dtype=torch.bfloat16)
                        items.append(((src.data_ptr() - self.w.data_ptr()) // 4, a, at, r, c))
                self._stage_table = ops.cast_transpose_table(self.w, items)
            ops.cast_transpose_bf16_batched(self.w, *self._stage_table)
        elif self.fp32_tensor_cores:
            # fp32 mode on the tensor cores: tf32 (hi, lo) parts of every dense kernel in both orientations
            for name, p in self.spec.params.items():
                leaf = name.split("/")[1]
                if leaf == "pointwise_kernel" or (leaf == "kernel" and p.shape[0] == 2):
                    src = self._mat(name)
                    r, c = src.shape
                    if name not in self._stage32:
                        mk = lambda *sh: torch.empty(sh, device=self.device, dtype=torch.float32)
                        self._stage32[name] = (src, mk(r, c))                  # the kernel itself is its own `hi` operand
                        self._stage32[name + "^T"] = (mk(c, r), mk(c, r))
                    ops.split_tf32(src, None, self._stage32[name][1])
                    ops.split_tf32(src, *self._stage32[name + "^T"], transpose=True)
        self._stage_dirty = False

    def _refold(self) -> None:
        if not self._fold_dirty or not self.use_bn:
            self._fold_dirty = False
            return
        for b in self.spec.blocks:
            o, c = self._bn_off[b.prefix]
            ops.bn_fold(self.wview(f"{b.prefix}_bn/gamma"), self.wview(f"{b.prefix}_bn/beta"),
                        self.wview(f"{b.prefix}_bn/moving_mean"), self.wview(f"{b.prefix}_bn/moving_variance"),
                        BN_EPS, self.fold[0, o:o + c], self.fold[1, o:o + c])
            if self.act_dtype == torch.bfloat16 and b.cin % 8 == 0:
                # inference: the BatchNormalization scale goes into the bf16 pointwise kernel ([Cout, Cin] staging), so the
                # GEMM epilogue only adds the shift (half the broadcast shared-memory loads per accumulator row)
                name = f"{b.prefix}_sepconv/pointwise_kernel"
                t = self._stage_fold.get(name)
                if t is None:
                    t = self._stage_fold[name] = torch.empty((b.cout, b.cin), device=self.device, dtype=torch.bfloat16)
                ops.cast_transpose_bf16(self._mat(name), None, t, col_scale=self.fold[0, o:o + c])
        self._fold_dirty = False

    # ------------------------------------------------------------------------------------------------ GEMM dispatch
    def _pw_fwd(self, prefix, d, out, **kw):
        name = f"{prefix}_sepconv/pointwise_kernel"
        cin = d.shape[-1]
        if self.act_dtype == torch.bfloat16:
            if cin % 8 == 0:
                ops.gemm(d, self._stage[name + "^T"], out, b_trans=True, **kw)       # tcgen05: B as [Cout, Cin]
            else:
                ops.gemm(d, self._stage[name], out, **kw)                            # K = 3: CUDA cores
        elif name + "^T" in self._stage32 and cin % 4 == 0:
            hi, lo = self._stage32[name + "^T"]                                      # tf32x3 on tcgen05: B as [Cout, Cin]
            ops.gemm(d, hi, out, b_trans=True, B_lo=lo, **kw)
        else:
            ops.gemm(d, self._mat(name), out, **kw)

    def _pw_dgrad(self, prefix, dz, dd, dz_lo=None):
        name = f"{prefix}_sepconv/pointwise_kernel"
        if self.act_dtype != torch.bfloat16 and name in self._stage32:
            hi, lo = self._stage32[name]                                             # [Cin, Cout] = [N, K]
            ops.gemm(dz, hi, dd, b_trans=True, B_lo=lo, A_lo=dz_lo)
            return
        B = self._stage[name] if self.act_dtype == torch.bfloat16 else self._mat(name)    # [Cin, Cout] = [N, K]
        ops.gemm(dz, B, dd, b_trans=True)

    def _convt_fwd(self, s, x, dst, drop):
        name = f"dec{s}_upsample/kernel"
        B = self._stage[name] if self.act_dtype == torch.bfloat16 else self._mat(name)    # [(a,b,co), Cin] = [N, K]
        B_lo = None
        if self.act_dtype != torch.bfloat16 and name in self._stage32:
            B, B_lo = self._stage32[name]
        ops.gemm(x, B, dst, b_trans=True, epilogue=ops.EPI_CONVT, shift=self.wview(f"dec{s}_upsample/bias"),
                 convt_hw=(x.shape[1], x.shape[2]), drop=drop, B_lo=B_lo)

    def _convt_dgrad(self, s, g2d, dx, g_lo=None):
        name = f"dec{s}_upsample/kernel"
        if self.act_dtype == torch.bfloat16:
            ops.gemm(g2d, self._stage[name + "^T"], dx, b_trans=True)                # B as [Cin, 4Cout] = [N, K]
        elif name + "^T" in self._stage32:
            hi, lo = self._stage32[name + "^T"]
            ops.gemm(g2d, hi, dx, b_trans=True, B_lo=lo, A_lo=g_lo)
        else:
            ops.gemm(g2d, self._mat(name), dx)                                       # B as [4Cout, Cin] = [K, N]

    # ------------------------------------------------------------------------------------------------ plans
    def _plan(self, batch: int, training: bool) -> _Plan:
        key = (batch, training)
        p = self._plans.get(key)
        if p is None:
            p = self._plans[key] = _Plan(self.device, self.act_dtype)
        return p

    def _dims(self, level: int) -> Tuple[int, int]:
        H, W, _ = self.spec.input_size
        return H >> level, W >> level

    def _drop(self, name: str, ctot: int, c0: int = 0):
        if self.dropout_rate <= 0.0:
            return None
        return ops.make_dropout(self.dropout_rate, self._drop_seed[name], ctot, c0,
                                self.step_word if self.dropout_masks_from_step else None)

    # ------------------------------------------------------------------------------------------------ inference
    def _infer_pw(self, prefix: str):
        """(bf16 [Cout, Cin] pointwise kernel, scale, shift) of an inference conv_block on the tensor-core path: with
        BatchNormalization the scale is already inside the kernel (scale None)."""
        name = f"{prefix}_sepconv/pointwise_kernel"
        o, c = self._bn_off[prefix]
        if not self.use_bn:
            return self._stage[name + "^T"], None, self.wview(f"{prefix}_sepconv/bias")
        if self.fold_scale_into_weights and name in self._stage_fold:
            return self._stage_fold[name], None, self.fold[1, o:o + c]
        return self._stage[name + "^T"], self.fold[0, o:o + c], self.fold[1, o:o + c]

    def _block_infer(self, pl: _Plan, prefix: str, x: torch.Tensor, y: torch.Tensor, pooled: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pooled: also produce MaxPooling2D((2,2)) of y (from the fused kernel's staged tile where that kernel runs)."""
        B, h, w, cin = x.shape
        if ops.stem_supported(cin, y.shape[-1]) and x.is_contiguous():
            o, c = self._bn_off[prefix]
            ops.stem_fwd(x, self._mat(f"{prefix}_sepconv/depthwise_kernel"), self._mat(f"{prefix}_sepconv/pointwise_kernel"), y,
                         scale=self.fold[0, o:o + c] if self.use_bn else None,
                         shift=self.fold[1, o:o + c] if self.use_bn else self.wview(f"{prefix}_sepconv/bias"), relu=True,
                         d_out=pl.buf(prefix + "/d3", (B, h, w, cin), torch.float32))
            return y
        if self.fuse_sepconv and ops.sepconv_fused_supported(x, y.shape[-1]):
            # whole conv_block in one kernel: the depthwise result is produced on chip as the GEMM's A operand
            wpt, sc, sh = self._infer_pw(prefix)
            fuse_pool = pooled is not None and self.fuse_pool and h % 2 == 0 and w % 2 == 0
            ops.sepconv_fused(x, self._mat(f"{prefix}_sepconv/depthwise_kernel"), wpt, y, scale=sc, shift=sh, relu=True,
                              pooled=pooled if fuse_pool else None)
            if pooled is not None and not fuse_pool:
                ops.maxpool2x2(y, pooled)
            return y
        level = (self.spec.input_size[0] // h).bit_length() - 1
        max_cin = 1024 if level == 4 else 2 * FILTERS[level]
        d = pl.buf(f"d{level}", (B * h * w * max_cin,))[: B * h * w * cin].view(B, h, w, cin)
        ops.dwconv3x3(x, self._mat(f"{prefix}_sepconv/depthwise_kernel"), d)
        if self.act_dtype == torch.bfloat16 and cin % 8 == 0:
            wpt, sc, sh = self._infer_pw(prefix)
            ops.gemm(d, wpt, y, b_trans=True, epilogue=ops.EPI_AFFINE_RELU, scale=sc, shift=sh)
        elif self.use_bn:
            o, c = self._bn_off[prefix]
            self._pw_fwd(prefix, d, y, epilogue=ops.EPI_AFFINE_RELU, scale=self.fold[0, o:o + c], shift=self.fold[1, o:o + c])
        else:
            self._pw_fwd(prefix, d, y, epilogue=ops.EPI_AFFINE_RELU, shift=self.wview(f"{prefix}_sepconv/bias"))
        if pooled is not None:
            ops.maxpool2x2(y, pooled)
        return y

    def forward_inference(self, x: torch.Tensor) -> torch.Tensor:
        """x: device fp32 [B,H,W,Cin] -> probabilities fp32 [B,H,W,classes] (plan-owned buffer, valid until the next call).
        With `use_graphs` the launch sequence of a batch size is captured once into a CUDA graph and replayed."""
        if not self.use_graphs or ops._prof is not None:
            return self._forward_inference_eager(x)
        B = x.shape[0]
        self._restage(); self._refold()              # parameter staging stays outside the graph (it is conditional)
        key = ("infer", B)
        ent = self._graphs.get(key)
        if ent is None:
            pl = self._plan(B, False)
            gx = pl.buf("graph_x", tuple(x.shape), torch.float32)
            gx.copy_(x)
            out = self._forward_inference_eager(gx)          # warm-up: allocates every buffer, sets kernel attributes
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._forward_inference_eager(gx)
            ent = self._graphs[key] = (g, gx, out)
        g, gx, out = ent
        if x.data_ptr() != gx.data_ptr():
            gx.copy_(x)
        g.replay()
        return out

    def _forward_inference_eager(self, x: torch.Tensor) -> torch.Tensor:
        B = x.shape[0]
        pl = self._plan(B, False)
        self._restage(); self._refold()
        H, W, Cin = self.spec.input_size
        if tuple(x.shape[1:]) != (H, W, Cin):
            raise ValueError(f"expected input (B,{H},{W},{Cin}), got {tuple(x.shape)}")
        if self.act_dtype == torch.bfloat16:
            cur = pl.buf("x_act", (B, H, W, Cin))
            ops.cast(x, cur)
        else:
            cur = x
        cats = {}
        for s in range(1, 5):
            f = FILTERS[s - 1]
            h, w = self._dims(s - 1)
            cat = cats[s] = pl.buf(f"cat{s}", (B, h, w, 2 * f))
            cur = self._block_infer(pl, f"enc{s}_block1", cur, pl.buf(f"ya{s}", (B, h, w, f)))
            nxt = pl.buf(f"pool{s}", (B, h // 2, w // 2, f))
            self._block_infer(pl, f"enc{s}_block2", cur, cat[..., f:], pooled=nxt)
            cur = nxt
        h, w = self._dims(4)
        cur = self._block_infer(pl, "bneck_block1", cur, pl.buf("ya5", (B, h, w, 1024)))
        cur = self._block_infer(pl, "bneck_block2", cur, pl.buf("yb5", (B, h, w, 1024)))
        for s in (4, 3, 2, 1):
            f = FILTERS[s - 1]
            h, w = self._dims(s - 1)
            self._convt_fwd(s, cur, cats[s][..., :f], None)
            cur = self._block_infer(pl, f"dec{s}_block1", cats[s], pl.buf(f"ya{s}", (B, h, w, f)))
            if s > 1 or not self.fuse_head or self.act_dtype != torch.bfloat16:
                cur = self._block_infer(pl, f"dec{s}_block2", cur, pl.buf(f"yb{s}", (B, h, w, f)))
        probs = pl.buf("probs", (B, H, W, self.num_classes), torch.float32)
        if self.fuse_head and self.act_dtype == torch.bfloat16:
            # dec1_block2's pointwise GEMM applies BN + ReLU and the 1x1 sigmoid/softmax head in its epilogue; the
            # 64-channel activation it would have produced is never written
            prefix = "dec1_block2"
            wpt, sc, sh = self._infer_pw(prefix)
            hw, hb = self._mat("output_mask/kernel"), self.wview("output_mask/bias")
            if self.fuse_sepconv and ops.sepconv_fused_supported(cur, FILTERS[0]):
                ops.sepconv_fused(cur, self._mat(f"{prefix}_sepconv/depthwise_kernel"), wpt, None, scale=sc, shift=sh,
                                  head_w=hw, head_b=hb, head_out=probs)
            else:
                d = pl.buf("d0", (B * H * W * 2 * FILTERS[0],))[: B * H * W * FILTERS[0]].view(B, H, W, FILTERS[0])
                ops.dwconv3x3(cur, self._mat(f"{prefix}_sepconv/depthwise_kernel"), d)
                ops.gemm(d, wpt, None, b_trans=True, epilogue=ops.EPI_HEAD, scale=sc, shift=sh, head_w=hw, head_b=hb, head_out=probs)
        else:
            ops.head_fwd(cur, self._mat("output_mask/kernel"), self.wview("output_mask/bias"), probs)
        return probs

    # ------------------------------------------------------------------------------------------------ training
    def _folds(self, prefix: str) -> bool:
        """BatchNormalization backward of this block is folded into its GEMMs (its output feeds *_block2's depthwise
        convolution directly, whose fused backward kernel delivers the ReLU-masked gradient and the two BN reductions)."""
        if not (self.fold_bn_bwd and self.fuse_dw_bwd and self.use_bn and self.act_dtype == torch.bfloat16
                and prefix in self._fz_off):
            return False
        _, cin, cout = self._fz_off[prefix]
        return ops.stem_supported(cin, cout) or cin % 8 == 0      # the depthwise strip kernel supplies colsum(d)

    def _fold_bufs(self, prefix: str):
        """(sums [2,Cout], sd [Cin], G [Cin,2Cout]) views of the per-step zeroed buffer; (coef, wab, bias) persistent."""
        o, cin, cout = self._fz_off[prefix]
        sums = self._fz[o:o + 2 * cout].view(2, cout)
        sd = self._fz[o + 2 * cout:o + 2 * cout + cin]
        G = self._fz[o + 2 * cout + cin:o + 2 * cout + cin + 2 * cin * cout].view(cin, 2 * cout)
        w = self._fold_w.get(prefix)
        if w is None:
            w = self._fold_w[prefix] = (torch.empty((3, cout), device=self.device),
                                        torch.empty((cin, 2 * cout), device=self.device, dtype=torch.bfloat16),
                                        torch.empty(cin, device=self.device))
        return sums, sd, G, w[0], w[1], w[2]

    def _bn(self, prefix):
        o, c = self._bn_off[prefix]
        return (self.bn_vec[0, o:o + c], self.bn_vec[1, o:o + c], self.bn_vec[2, o:o + c], self.bn_vec[3, o:o + c])

    def _act_fused(self, prefix: str) -> bool:
        """The BN+ReLU pass of this *_block1 is skipped: its only consumers, *_block2's depthwise forward and fused
        depthwise backward kernels, form y = max(z*scale + shift, 0) on load."""
        return (self.fuse_bn_act and self.fuse_dw_bwd and prefix.endswith("_block1") and self._bn_off[prefix][1] % 8 == 0)

    def _act_affine(self, prefix: str):
        """(scale, shift) of a block's BatchNormalization + ReLU as applied to its pre-activation tensor."""
        o, c = self._bn_off[prefix]
        return (self.bn_vec[0, o:o + c], self.bn_vec[1, o:o + c]) if self.use_bn else (self.ones[:c], self.zeros[:c])

    def _b1_fwd(self, pl, prefix, x, yshape, xaff):
        """Forward of a *_block1; with `fuse_bn_act` its activation is never materialised and the pre-BN tensor is returned."""
        if self._act_fused(prefix):
            z = self._block_train_fwd(pl, prefix, x, None, defer_act=True)
            xaff[prefix[:-1] + "2"] = self._act_affine(prefix)
            return z
        return self._block_train_fwd(pl, prefix, x, pl.buf(prefix + "/y", yshape))

    def _block_train_fwd(self, pl, prefix, x, y, pooled=None, drop=None, x_affine=None, defer_act=False):
        """x_affine: x is the producer's pre-BN tensor; apply (scale, shift) + ReLU on load.  defer_act: do not run this
        block's own BN+ReLU pass (the consumer does it on load) and return the pre-BN tensor."""
        B, h, w, cin = x.shape
        o, c = self._bn_off[prefix]
        cout = c
        z = pl.buf(prefix + "/z", (B, h, w, cout))
        stem = ops.stem_supported(cin, cout) and x.is_contiguous()
        wd, wp = self._mat(f"{prefix}_sepconv/depthwise_kernel"), self._mat(f"{prefix}_sepconv/pointwise_kernel")
        if not stem:
            d = pl.buf(prefix + "/d", (B, h, w, cin))
            ops.dwconv3x3(x, wd, d, colsum=self._fold_bufs(prefix)[1] if self._folds(prefix) else None,
                          in_scale=x_affine[0] if x_affine else None, in_shift=x_affine[1] if x_affine else None)
        if self.use_bn:
            scale, shift, smean, srstd = self._bn(prefix)
            if stem:
                d3 = pl.buf(prefix + "/d3", (B, h, w, cin), torch.float32)
                ops.stem_fwd(x, wd, wp, z, colsum=self.colstats[0, o:o + c], colsq=self.colstats[1, o:o + c], d_out=d3)
            else:
                self._pw_fwd(prefix, d, z, epilogue=ops.EPI_STATS, colsum=self.colstats[0, o:o + c], colsq=self.colstats[1, o:o + c])
            ops.bn_finalize(self.colstats[0, o:o + c], self.colstats[1, o:o + c], B * h * w,
                            self.wview(f"{prefix}_bn/gamma"), self.wview(f"{prefix}_bn/beta"), BN_EPS, BN_MOMENTUM,
                            self.wview(f"{prefix}_bn/moving_mean"), self.wview(f"{prefix}_bn/moving_variance"),
                            scale, shift, smean, srstd)
        else:
            if stem:
                ops.stem_fwd(x, wd, wp, z, shift=self.wview(f"{prefix}_sepconv/bias"))
            else:
                self._pw_fwd(prefix, d, z, epilogue=ops.EPI_AFFINE, shift=self.wview(f"{prefix}_sepconv/bias"))
            scale, shift = self.ones[:c], self.zeros[:c]
        if defer_act:
            return z
        ops.bn_act(z, scale, shift, y, relu=True, pooled=pooled, drop=drop)
        return y

    def _block_train_bwd(self, pl, prefix, x, dy, scr, dx_out=None, ydrop=None, dx_drop=None, mask_for=None, folded=False,
                         dx_drop_from=0, x_affine=None, up_out=None):
        """dy: gradient w.r.t. the block output (as stored).  scr: two scratch tensors (flat).  Returns dx_out.
        mask_for: prefix of the block that produced x (= its post-ReLU output): dx_out then is the ReLU-masked gradient
        w.r.t. that block's BatchNormalization output and its two BN-backward reductions are accumulated on the way.
        folded: dy arrived that way, and this block's BatchNormalization backward is folded into the GEMMs below."""
        B, h, w, cin = x.shape
        z = pl.t[prefix + "/z"]
        cout = z.shape[-1]
        M = B * h * w
        stem = (prefix + "/d") not in pl.t
        dz = scr[0][: M * cout].view(B, h, w, cout)
        dd = scr[1][: M * cin].view(B, h, w, cin)
        if self.use_bn:
            scale, shift, smean, srstd = self._bn(prefix)
            dgamma, dbeta = self.wview(f"{prefix}_bn/gamma", self.g), self.wview(f"{prefix}_bn/beta", self.g)
        else:
            o, c = self._bn_off[prefix]
            scale, shift, smean, srstd = self.ones[:c], self.zeros[:c], None, None
            dgamma, dbeta = None, self.wview(f"{prefix}_sepconv/bias", self.g)
        gwp = self._mat(f"{prefix}_sepconv/pointwise_kernel", self.g)
        if folded:
            sums, sd, G, coef, wab, bias = self._fold_bufs(prefix)
            gamma, beta = self.wview(f"{prefix}_bn/gamma"), self.wview(f"{prefix}_bn/beta")
            if stem:   # streaming first-block backward: dz = A*g + B*z + K in registers, then a 3-channel depthwise weight gradient
                ops.bn_bwd_coef(sums, gamma, beta, smean, srstd, M, dgamma, dbeta, coef, guard=self.fold_guard)
                dd3 = pl.buf(prefix + "/dd3", (B, h, w, cin))
                ops.stem_bwd_folded(dy, z, coef, pl.t[prefix + "/d3"], self._mat(f"{prefix}_sepconv/pointwise_kernel"), gwp, dd3)
                ops.dwconv3x3_bwd_weight(x, dd3, self._mat(f"{prefix}_sepconv/depthwise_kernel", self.g))
                return None
            else:      # dz = A*g + B*z + K never exists: both contractions read [g | z]
                ops.bn_bwd_coef(sums, gamma, beta, smean, srstd, M, dgamma, dbeta, coef,
                                w=self._mat(f"{prefix}_sepconv/pointwise_kernel"), wab=wab, bias=bias, guard=self.fold_guard)
                d = pl.t[prefix + "/d"]
                if self.fuse_pw_bwd and ops.pw_bwd_fused_supported(dy, z, d, dd):
                    ops.pw_bwd_fused(dy, z, d, wab, bias, dd, G)    # both contractions from one pass over [g | z] and d
                    ops.bn_bwd_wgrad_combine(G, coef, sd, gwp)
                else:
                    ops.gemm(d, dy, G, a_trans=True, accumulate=True, B2=z)
                    ops.bn_bwd_wgrad_combine(G, coef, sd, gwp)
                    ops.gemm(dy, wab, dd, b_trans=True, A2=z, epilogue=ops.EPI_AFFINE, shift=bias)
        else:
            ops.bn_bwd_reduce(dy, z, scale, shift, smean, srstd, dgamma, dbeta, relu=True, drop=ydrop)
            ops.bn_bwd_apply(dy, z, scale, shift, smean, srstd, dgamma, dbeta, dz, relu=True, drop=ydrop)
        if stem:       # fused first block: both weight gradients from one pass over dz; the image needs no gradient
            ops.stem_bwd(x, dz, self._mat(f"{prefix}_sepconv/depthwise_kernel"), self._mat(f"{prefix}_sepconv/pointwise_kernel"),
                         self._mat(f"{prefix}_sepconv/depthwise_kernel", self.g), gwp)
            return None
        if not folded:
            d = pl.t[prefix + "/d"]
            # fp32 mode on the tensor cores: dz feeds both GEMMs, its tf32 `lo` part is made once
            dz_lo = ops.tf32_lo(dz, slot=2) if (self._stage32 and dz.dtype == torch.float32 and cout % 8 == 0) else None
            ops.gemm(d, dz, gwp, a_trans=True, accumulate=True, tf32x3=self.fp32_tensor_cores, B_lo=dz_lo)
            self._pw_dgrad(prefix, dz, dd, dz_lo=dz_lo)
        wd, gwd = self._mat(f"{prefix}_sepconv/depthwise_kernel"), self._mat(f"{prefix}_sepconv/depthwise_kernel", self.g)
        if dx_out is not None and self.fuse_dw_bwd and ops.dwconv3x3_bwd_supported(x, dd, dx_out):
            # both gradients from one pass over dd (+ the producer's ReLU mask and BN-backward reductions)
            ops.dwconv3x3_bwd(x, dd, wd, dx_out, gwd, drop=dx_drop, drop_c_from=dx_drop_from, relu_mask=mask_for is not None,
                              bn_sums=self._fold_bufs(mask_for)[0] if mask_for is not None else None,
                              x_scale=x_affine[0] if x_affine else None, x_shift=x_affine[1] if x_affine else None,
                              up_out=up_out[0] if up_out else None, up_colsum=up_out[1] if up_out else None)
            return dx_out
        if mask_for is not None or x_affine is not None or up_out is not None:
            raise RuntimeError(f"{prefix}: folded BN backward / BN+ReLU on load need the fused depthwise backward kernel")
        ops.dwconv3x3_bwd_weight(x, dd, gwd)
        if dx_out is not None:
            ops.dwconv3x3(dd, wd, dx_out, flip=True, drop=dx_drop)
        return dx_out

    def train_forward_backward(self, x: torch.Tensor, y_true: torch.Tensor, loss: str = "dice") -> torch.Tensor:
        """One training-mode forward + backward on device tensors (x fp32 [B,H,W,Cin], y_true fp32 [B,H,W,classes]).
        Leaves d loss / d theta in self.g and updates the BN moving statistics.  Returns out3 = (loss, dice_coef,
        iou_coef) as a device tensor (no host synchronisation)."""
        kind = {"dice": 0, "iou": 1}[loss]
        B = x.shape[0]
        H, W, Cin = self.spec.input_size
        NC = self.num_classes
        if tuple(x.shape) != (B, H, W, Cin) or tuple(y_true.shape) != (B, H, W, NC):
            raise ValueError(f"expected x (B,{H},{W},{Cin}) and y_true (B,{H},{W},{NC})")
        pl = self._plan(B, True)
        self._restage()
        self.g.zero_()
        self.colstats.zero_()
        self._fz.zero_()
        if self.act_dtype == torch.bfloat16:
            x0 = pl.buf("x_act", (B, H, W, Cin))
            ops.cast(x, x0)
        else:
            x0 = x
        # ---------------- forward
        cur = x0
        cats, xin = {}, {}
        xaff: Dict[str, tuple] = {}       # *_block2 -> (scale, shift) of *_block1 when its BN+ReLU is applied on load
        head_aff = None                   # (scale, shift) of dec1_block2 when the head kernels apply its BN+ReLU on load
        for s in range(1, 5):
            f = FILTERS[s - 1]
            h, w = self._dims(s - 1)
            cat = cats[s] = pl.buf(f"cat{s}", (B, h, w, 2 * f))
            xin[f"enc{s}_block1"] = cur
            y1 = self._b1_fwd(pl, f"enc{s}_block1", cur, (B, h, w, f), xaff)
            xin[f"enc{s}_block2"] = y1
            pooled = pl.buf(f"pool{s}", (B, h // 2, w // 2, f))
            sdrop = self._drop(f"dec{s}_dropout", 2 * f, f) if s > 1 else None
            self._block_train_fwd(pl, f"enc{s}_block2", y1, cat[..., f:], pooled=pooled, drop=sdrop, x_affine=xaff.get(f"enc{s}_block2"))
            cur = pooled
        h, w = self._dims(4)
        xin["bneck_block1"] = cur
        y1 = self._b1_fwd(pl, "bneck_block1", cur, (B, h, w, 1024), xaff)
        xin["bneck_block2"] = y1
        cur = self._block_train_fwd(pl, "bneck_block2", y1, pl.buf("bneck_block2/y", (B, h, w, 1024)),
                                    drop=self._drop("bneck_dropout", 1024), x_affine=xaff.get("bneck_block2"))
        convt_in = {}
        for s in (4, 3, 2, 1):
            f = FILTERS[s - 1]
            h, w = self._dims(s - 1)
            convt_in[s] = cur
            self._convt_fwd(s, cur, cats[s][..., :f], self._drop(f"dec{s}_dropout", 2 * f, 0) if s > 1 else None)
            xin[f"dec{s}_block1"] = cats[s]
            y1 = self._b1_fwd(pl, f"dec{s}_block1", cats[s], (B, h, w, f), xaff)
            xin[f"dec{s}_block2"] = y1
            if s == 1 and self.fuse_bn_act and self.act_dtype == torch.bfloat16 and NC == 1 and f == 64:
                # dec1_block2's BN+ReLU is applied on load by the streamed head kernels (forward and backward)
                cur = self._block_train_fwd(pl, "dec1_block2", y1, None, x_affine=xaff.get("dec1_block2"), defer_act=True)
                head_aff = self._act_affine("dec1_block2")
            else:
                cur = self._block_train_fwd(pl, f"dec{s}_block2", y1, pl.buf(f"dec{s}_block2/y", (B, h, w, f)),
                                            x_affine=xaff.get(f"dec{s}_block2"))
        probs = pl.buf("probs", (B, H, W, NC), torch.float32)
        sums = pl.buf("sums", (B, NC, 3), torch.float64)
        sums.zero_()
        wk, bk = self._mat("output_mask/kernel"), self.wview("output_mask/bias")
        ops.head_fwd(cur, wk, bk, probs, y_true, sums, x_scale=head_aff[0] if head_aff else None,
                     x_shift=head_aff[1] if head_aff else None)
        out3 = pl.buf("out3", (3,), torch.float32)
        coef = pl.buf("coef", (B, NC, 2), torch.float32)
        ops.seg_loss_finalize(sums, B * NC, SMOOTH, kind, 1.0, out3, coef)
        # ---------------- backward
        n_scr = B * H * W * 128
        S = [pl.buf(f"scr{i}", (n_scr,)) for i in range(3)]
        dy = S[0][: B * H * W * 64].view(B, H, W, 64)
        fold_d1 = self._folds("dec1_block2")
        ops.head_bwd(cur, wk, probs, y_true, coef, dy, self._mat("output_mask/kernel", self.g),
                     self.wview("output_mask/bias", self.g), bn_sums=self._fold_bufs("dec1_block2")[0] if fold_d1 else None,
                     x_scale=head_aff[0] if head_aff else None, x_shift=head_aff[1] if head_aff else None)
        ci = 0   # index of the scratch buffer that currently holds dy
        dcat = {}
        for s in (1, 2, 3, 4):
            f = FILTERS[s - 1]
            h, w = self._dims(s - 1)
            M = B * h * w
            o1, o2 = (ci + 1) % 3, (ci + 2) % 3
            # dec{s}_block2: dy (S[ci]) -> dx into S[ci] (dy is dead once dz exists)
            dx = S[ci][: M * f].view(B, h, w, f)
            fold = self._folds(f"dec{s}_block1")
            self._block_train_bwd(pl, f"dec{s}_block2", xin[f"dec{s}_block2"], dy, (S[o1], S[o2]), dx_out=dx,
                                  mask_for=f"dec{s}_block1" if fold else None, folded=(s == 1 and fold_d1),
                                  x_affine=xaff.get(f"dec{s}_block2"))
            # dec{s}_block1: input is the (dropped-out) concat buffer
            dcat[s] = pl.buf(f"dcat{s}", (B, h, w, 2 * f))
            # the Dropout mask of the upsampled half of the concat gradient is applied by its reader (the un-pixel-shuffle
            # gather, a memory-bound kernel with ALU to spare); the depthwise backward kernel masks only the skip half
            xi = convt_in[s]
            Mi = xi.shape[0] * xi.shape[1] * xi.shape[2]
            gth = S[o1][: Mi * 4 * f].view(Mi, 4 * f)            # S[o1] held dz, dead once dd exists
            dbias = self.wview(f"dec{s}_upsample/bias", self.g)
            # gather-free: the depthwise backward kernel stores the upsampled half of the concat gradient un-pixel-shuffled
            # (the operand of the Conv2DTranspose gradient GEMMs) and accumulates the bias gradient.  Only where the concat
            # carries no Dropout (dec1, u_net.py:97: `i < len(filters)-1`): with Dropout the mask of the upsampled half is
            # cheaper in the memory-bound gather than in the issue-bound depthwise kernel (measured: +0.46 vs -0.44 ms).
            cat_drop = self._drop(f"dec{s}_dropout", 2 * f, 0) if s > 1 else None
            direct = (self.convt_bwd_direct and self.fuse_dw_bwd and (cat_drop is None or self.convt_bwd_direct_drop)
                      and f % (64 if self.act_dtype == torch.bfloat16 else 32) == 0
                      and ops.dwconv3x3_bwd_supported(cats[s], dx, dcat[s]))
            defer = (not direct) and self.defer_dropout and s > 1 and f % 64 == 0
            self._block_train_bwd(pl, f"dec{s}_block1", cats[s], dx, (S[o1], S[o2]), dx_out=dcat[s],
                                  dx_drop=self._drop(f"dec{s}_dropout", 2 * f, 0) if s > 1 else None,
                                  dx_drop_from=f if defer else 0, folded=fold, up_out=(gth, dbias) if direct else None)
            # Conv2DTranspose backward
            if not direct:
                ops.convt_bwd_gather(dcat[s][..., :f], gth, dbias,
                                     drop=self._drop(f"dec{s}_dropout", 2 * f, 0) if defer else None)
            g_lo = ops.tf32_lo(gth, slot=2) if (self._stage32 and gth.dtype == torch.float32) else None
            ops.gemm(gth, xi, self._mat(f"dec{s}_upsample/kernel", self.g), a_trans=True, accumulate=True, tf32x3=self.fp32_tensor_cores,
                     A_lo=g_lo)
            dy = S[ci][: Mi * 2 * f].view(xi.shape)
            self._convt_dgrad(s, gth, dy, g_lo=g_lo)
        if self.grad_hook:
            self.grad_hook("decoder")
        # bottleneck
        h, w = self._dims(4)
        M = B * h * w
        o1, o2 = (ci + 1) % 3, (ci + 2) % 3
        dx = S[ci][: M * 1024].view(B, h, w, 1024)
        fold = self._folds("bneck_block1")
        self._block_train_bwd(pl, "bneck_block2", xin["bneck_block2"], dy, (S[o1], S[o2]), dx_out=dx,
                              ydrop=self._drop("bneck_dropout", 1024), mask_for="bneck_block1" if fold else None,
                              x_affine=xaff.get("bneck_block2"))
        dpool = S[ci][: M * 512].view(B, h, w, 512)
        self._block_train_bwd(pl, "bneck_block1", xin["bneck_block1"], dx, (S[o1], S[o2]), dx_out=dpool, folded=fold)
        if self.grad_hook:
            self.grad_hook("bottleneck")
        # encoder
        for s in (4, 3, 2, 1):
            f = FILTERS[s - 1]
            h, w = self._dims(s - 1)
            M = B * h * w
            o1, o2 = (ci + 1) % 3, (ci + 2) % 3
            if self.use_bn:
                scale, shift, _, _ = self._bn(f"enc{s}_block2")
            else:
                scale, shift = self.ones[:f], self.zeros[:f]
            dy = S[o1][: M * f].view(B, h, w, f)
            fold2 = self._folds(f"enc{s}_block2")
            ops.maxpool2x2_bwd(pl.t[f"enc{s}_block2/z"], scale, shift, dpool, dcat[s][..., f:], dy,
                               bn_sums=self._fold_bufs(f"enc{s}_block2")[0] if fold2 else None)
            ci, o1 = o1, ci        # dy now lives in the old o1; the old ci is free
            dx = S[ci][: M * f].view(B, h, w, f)
            fold = self._folds(f"enc{s}_block1")
            self._block_train_bwd(pl, f"enc{s}_block2", xin[f"enc{s}_block2"], dy, (S[o1], S[o2]), dx_out=dx,
                                  mask_for=f"enc{s}_block1" if fold else None, folded=fold2, x_affine=xaff.get(f"enc{s}_block2"))
            x1 = xin[f"enc{s}_block1"]
            cin = x1.shape[-1]
            dpool = S[ci][: M * cin].view(B, h, w, cin) if s > 1 else None
            self._block_train_bwd(pl, f"enc{s}_block1", x1, dx, (S[o1], S[o2]), dx_out=dpool, folded=fold)
        if self.grad_hook:
            self.grad_hook("encoder")
        self._fold_dirty = True     # moving statistics changed
        return out3

    def apply_gradients(self) -> None:
        """Keras-form AdamW over the whole flat parameter buffer (train.py:226), then advance t and the dropout word."""
        ops.adamw_step(self.w, self.g, self.m, self.v, self.hyper)
        ops.step_advance(self.hyper, self.step_word)
        self._stage_dirty = True
        self._restage()

    def _step_eager(self, x, y_true, loss):
        out3 = self.train_forward_backward(x, y_true, loss)
        if self.grad_finish is not None:
            self.grad_finish()
        self.apply_gradients()
        return out3

    def train_step(self, x: torch.Tensor, y_true: torch.Tensor, loss: str = "dice") -> torch.Tensor:
        """forward + backward (+ data-parallel gradient exchange) + AdamW.  Steps are replayed from a CUDA graph when
        `use_graphs` is set (the learning rate, the Adam step count and the dropout seed word live in device memory, so a
        replay is a new step); under data parallel the NCCL exchanges are part of the graph (`graph_collectives`)."""
        if (not self.use_graphs or ops._prof is not None
                or (self.grad_hook is not None and not self.graph_collectives)):
            return self._step_eager(x, y_true, loss)
        B = x.shape[0]
        key = ("train", B, loss)
        ent = self._graphs.get(key)
        if ent is None:
            pl = self._plan(B, True)
            gx = pl.buf("graph_x", tuple(x.shape), torch.float32)
            gy = pl.buf("graph_y", tuple(y_true.shape), torch.float32)
            gx.copy_(x); gy.copy_(y_true)
            out3 = self._step_eager(gx, gy, loss)                # a real (eager) step doubles as the warm-up
            torch.cuda.current_stream().synchronize()
            g = torch.cuda.CUDAGraph()
            try:
                # data parallel: NCCL's watchdog thread polls CUDA events concurrently, which the default (global) capture
                # mode would turn into a capture error; only this thread's calls belong to the capture
                mode = "thread_local" if self.grad_hook is not None else "global"
                with torch.cuda.graph(g, capture_error_mode=mode):
                    cap = self._step_eager(gx, gy, loss)
            except Exception as e:
                if self.grad_hook is None:
                    raise
                # collectives could not be captured on this NCCL / driver combination: keep launching eagerly
                import warnings
                warnings.warn(f"CUDA-graph capture of the data-parallel step failed ({e}); running eagerly")
                self.graph_collectives = False
                torch.cuda.synchronize()
                return out3
            self._graphs[key] = (g, gx, gy, cap)
            return out3
        g, gx, gy, out3 = ent
        self._restage()          # set_weights / load_weights since the last step: refresh the bf16 operand copies (conditional,
                                 # so it lives outside the graph; the graph itself restages after its own AdamW step)
        if x.data_ptr() != gx.data_ptr():
            gx.copy_(x)
        if y_true.data_ptr() != gy.data_ptr():
            gy.copy_(y_true)
        g.replay()
        self._fold_dirty = True
        return out3

    # ------------------------------------------------------------------------------------------------ evaluation
    def evaluate_batch(self, x: torch.Tensor, y_true: torch.Tensor, loss: str = "dice") -> torch.Tensor:
        """Inference-mode forward + (loss, dice_coef, iou_coef) on device."""
        probs = self.forward_inference(x)
        pl = self._plan(x.shape[0], False)
        B, NC = x.shape[0], self.num_classes
        sums = pl.buf("sums", (B, NC, 3), torch.float64)
        sums.zero_()
        ops.seg_sums(y_true, probs, sums)
        out3 = pl.buf("out3", (3,), torch.float32)
        ops.seg_loss_finalize(sums, B * NC, SMOOTH, {"dice": 0, "iou": 1}[loss], 1.0, out3, None)
        return out3

    def check_fold_guard(self) -> bool:
        """Host read of the conditioning flag of the folded BatchNormalization backward (call at a point that synchronises
        anyway, e.g. the end of an epoch).  When some |gamma| has fallen below |beta|/16 the engine switches to the explicit
        reduce / apply schedule — which forms sum(g*xhat) from z directly — for the rest of training.  Returns True if it did."""
        if not self.fold_bn_bwd or int(self.fold_guard.item()) == 0:
            return False
        import warnings
        warnings.warn("BatchNormalization backward: a gamma fell below |beta|/16; switching from the folded to the two-pass schedule")
        self.fold_bn_bwd = False
        self.fold_guard.zero_()
        self._graphs = {k: v for k, v in self._graphs.items() if k[0] != "train"}
        return True

    def plan_bytes(self) -> int:
        return sum(p.bytes() for p in self._plans.values())

    def release_plans(self) -> None:
        self._graphs.clear()
        self._plans.clear()
