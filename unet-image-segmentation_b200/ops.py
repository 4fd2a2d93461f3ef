"""Typed wrappers over the C-ABI (include/unet_b200.h): torch tensors in, kernel launches on the current stream out.

torch is used for device memory and streams only; every function here ends in exactly one call into
libunet_b200.so.  Activations are NHWC tensors or channel-slice views of wider NHWC buffers (`x[..., c0:c0+C]`):
the leading dimension handed to the kernels is the pixel stride of the view, which is how Concatenate
(reference model/u_net.py:96) costs nothing.  Nothing in this module falls back to torch arithmetic.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (EPI_AFFINE, EPI_AFFINE_RELU, EPI_CONVT, EPI_HEAD, EPI_NONE, EPI_STATS, UNET_BF16, UNET_F32, Dropout,
                   GemmArgs)

_DT = {torch.float32: UNET_F32, torch.bfloat16: UNET_BF16}

launches = 0   # number of kernel-launching C-ABI calls made through this module (bench.py reports it)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _dt(t: torch.Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise TypeError(f"unsupported activation dtype {t.dtype}") from None


def _f32(t: Optional[torch.Tensor], name: str) -> None:
    if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or not t.is_cuda):
        raise TypeError(f"{name} must be a contiguous CUDA fp32 tensor")


def _nhwc(t: torch.Tensor, name: str) -> Tuple[int, int, int, int, int]:
    """(N, H, W, C, ld) of an NHWC tensor or channel-slice view."""
    if t.dim() != 4 or not t.is_cuda:
        raise ValueError(f"{name}: expected a 4-D CUDA tensor (N,H,W,C)")
    n, h, w, c = t.shape
    ld = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if n > 1 else c))
    if c > 1 and t.stride(3) != 1:
        raise ValueError(f"{name}: channels must be contiguous")
    if (w > 1 and t.stride(2) != ld) or (h > 1 and t.stride(1) != w * ld) or (n > 1 and t.stride(0) != h * w * ld):
        raise ValueError(f"{name}: not a pixel-strided NHWC view (strides {t.stride()})")
    return n, h, w, c, ld


def _rows(t: torch.Tensor, name: str) -> Tuple[int, int, int]:
    """(rows, cols, ld) of a 2-D row-major matrix or column-slice view; NHWC tensors are flattened over pixels."""
    if t.dim() == 4:
        n, h, w, c, ld = _nhwc(t, name)
        return n * h * w, c, ld
    if t.dim() != 2 or not t.is_cuda:
        raise ValueError(f"{name}: expected a 2-D CUDA tensor")
    if t.shape[1] > 1 and t.stride(1) != 1:
        raise ValueError(f"{name}: columns must be contiguous")
    return t.shape[0], t.shape[1], (t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0)))


def make_dropout(rate: float, seed: int, ctot: int, c0: int = 0,
                 seed_dev: Optional[torch.Tensor] = None) -> Optional[Dropout]:
    """Dropout(rate) mask descriptor (u_net.py:78,98).  `seed_dev`: int32 device word added to `seed` in-kernel."""
    if rate <= 0.0:
        return None
    if seed_dev is not None and (seed_dev.dtype != torch.int32 or not seed_dev.is_cuda):
        raise TypeError("seed_dev must be a CUDA int32 tensor")
    return Dropout(float(rate), int(seed) & 0xFFFFFFFF, int(ctot), int(c0), _p(seed_dev))


def _dref(d: Optional[Dropout]):
    return None if d is None else C.byref(d)


# Optional per-launch timing (bench.py): CUDA events recorded on the launching stream around every C-ABI call,
# keyed by (entry point, shape tag), together with the call's ALGORITHMIC bytes and flops (each distinct input read
# once + each output written once; 2*M*N*K for contractions) so that a roofline fraction can be formed per kernel.
_prof = None
_tag = ""


def profile_begin() -> None:
    global _prof
    _prof = {}


def profile_end() -> dict:
    """-> {key: dict(calls, ms, bytes, flops)}; bytes/flops are per-call algorithmic figures summed over calls."""
    global _prof
    torch.cuda.synchronize()
    out = {}
    for key, rec in (_prof or {}).items():
        ms = sum(a.elapsed_time(b) for a, b in rec["ev"])
        out[key] = dict(calls=len(rec["ev"]), ms=ms, bytes=rec["bytes"], flops=rec["flops"])
    _prof = None
    return out


def _nbytes(*tensors) -> int:
    return sum(t.numel() * t.element_size() for t in tensors if t is not None)


def _call(name: str, *args, tag: str = "", nbytes: int = 0, flops: int = 0) -> None:
    global launches
    launches += 1
    if _prof is None:
        _lib.call(name, *args)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.call(name, *args)
    e1.record()
    rec = _prof.setdefault(f"{name[5:]}[{tag}]", dict(ev=[], bytes=0, flops=0))
    rec["ev"].append((e0, e1)); rec["bytes"] += nbytes; rec["flops"] += flops


# ------------------------------------------------------------------------------------------------ depthwise
def dwconv3x3(x: torch.Tensor, w9c: torch.Tensor, y: torch.Tensor, flip: bool = False,
              in_scale: Optional[torch.Tensor] = None, in_shift: Optional[torch.Tensor] = None,
              drop: Optional[Dropout] = None, colsum: Optional[torch.Tensor] = None) -> None:
    """SeparableConv2D depthwise half (u_net.py:14-20); flip=True gives the gradient w.r.t. the input.
    colsum (fp32 [C], accumulated): per-channel sums of the stored output."""
    n, h, w, c, ldx = _nhwc(x, "x")
    n2, h2, w2, c2, ldy = _nhwc(y, "y")
    if (n, h, w, c) != (n2, h2, w2, c2) or x.dtype != y.dtype:
        raise ValueError("dwconv3x3: x and y disagree")
    _f32(w9c, "w9c"); _f32(in_scale, "in_scale"); _f32(in_shift, "in_shift"); _f32(colsum, "colsum")
    if w9c.numel() != 9 * c or (colsum is not None and colsum.numel() != c):
        raise ValueError("dwconv3x3: w9c must hold 9*C floats (and colsum C)")
    _call("unet_dwconv3x3_fwd", _p(x), ldx, _p(w9c), _p(y), ldy, n, h, w, c, _dt(x), int(flip),
          _p(in_scale), _p(in_shift), _dref(drop), _p(colsum), _stream(),
          tag=f"{n}x{h}x{w}x{c}", nbytes=_nbytes(x, y, w9c), flops=18 * x.numel())


def dwconv3x3_bwd_weight(x: torch.Tensor, dy: torch.Tensor, dw9c: torch.Tensor) -> None:
    n, h, w, c, ldx = _nhwc(x, "x")
    n2, h2, w2, c2, lddy = _nhwc(dy, "dy")
    if (n, h, w, c) != (n2, h2, w2, c2) or x.dtype != dy.dtype:
        raise ValueError("dwconv3x3_bwd_weight: x and dy disagree")
    _f32(dw9c, "dw9c")
    _call("unet_dwconv3x3_bwd_weight", _p(x), ldx, _p(dy), lddy, _p(dw9c), n, h, w, c, _dt(x), _stream(),
          tag=f"{n}x{h}x{w}x{c}", nbytes=_nbytes(x, dy, dw9c), flops=18 * x.numel())


def dwconv3x3_bwd_supported(x: torch.Tensor, dy: torch.Tensor, dx: torch.Tensor) -> bool:
    nv = 8 // x.element_size()
    c = x.shape[-1]
    return (c % nv == 0 and c >= 8 and all(t.data_ptr() % 16 == 0 and (t.stride(-2) * t.element_size()) % 16 == 0
                                          for t in (x, dy, dx)))


def dwconv3x3_bwd(x: torch.Tensor, dy: torch.Tensor, w9c: torch.Tensor, dx: torch.Tensor, dw9c: torch.Tensor,
                  relu_mask: bool = False, bn_sums: Optional[torch.Tensor] = None, drop: Optional[Dropout] = None,
                  drop_c_from: int = 0, x_scale: Optional[torch.Tensor] = None, x_shift: Optional[torch.Tensor] = None,
                  up_out: Optional[torch.Tensor] = None, up_colsum: Optional[torch.Tensor] = None) -> None:
    """SeparableConv2D depthwise backward in one pass over dy: input gradient (optionally ReLU-masked by x > 0, with the
    BatchNormalization-backward reductions sum(g), sum(g*x) accumulated into bn_sums [2,C]) + weight gradient.
    up_out [N*H/2*W/2, 4*up_c] (contiguous): the first up_c channels of dx are stored un-pixel-shuffled there (the operand of
    the Conv2DTranspose gradient GEMMs) instead of into dx; up_colsum (fp32 [up_c], accumulated): their per-channel sums."""
    n, h, w, c, ldx = _nhwc(x, "x")
    n2, h2, w2, c2, lddy = _nhwc(dy, "dy")
    n3, h3, w3, c3, lddx = _nhwc(dx, "dx")
    if (n, h, w, c) != (n2, h2, w2, c2) or (n, h, w, c) != (n3, h3, w3, c3) or x.dtype != dy.dtype or x.dtype != dx.dtype:
        raise ValueError("dwconv3x3_bwd: x, dy and dx disagree")
    _f32(w9c, "w9c"); _f32(dw9c, "dw9c"); _f32(bn_sums, "bn_sums"); _f32(x_scale, "x_scale"); _f32(x_shift, "x_shift")
    if w9c.numel() != 9 * c or dw9c.numel() != 9 * c or (bn_sums is not None and bn_sums.numel() != 2 * c):
        raise ValueError("dwconv3x3_bwd: w9c / dw9c must hold 9*C floats and bn_sums 2*C")
    up_c = 0
    if up_out is not None:
        if up_out.dtype != x.dtype or up_out.dim() != 2 or not up_out.is_contiguous() or up_out.shape[1] % 4 != 0 \
                or up_out.shape[0] * 4 != n * h * w:
            raise ValueError("dwconv3x3_bwd: up_out must be a contiguous [N*H/2*W/2, 4*up_c] tensor of x's dtype")
        up_c = up_out.shape[1] // 4
        _f32(up_colsum, "up_colsum")
        if up_colsum is not None and up_colsum.numel() != up_c:
            raise ValueError("dwconv3x3_bwd: up_colsum must hold up_c floats")
    elif up_colsum is not None:
        raise ValueError("dwconv3x3_bwd: up_colsum needs up_out")
    _call("unet_dwconv3x3_bwd", _p(x), ldx, _p(dy), lddy, _p(w9c), _p(dx), lddx, _p(dw9c), n, h, w, c, _dt(x),
          int(relu_mask), _p(bn_sums), _dref(drop), int(drop_c_from), _p(x_scale), _p(x_shift),
          _p(up_out), up_c, _p(up_colsum), _stream(),
          tag=f"{n}x{h}x{w}x{c}" + ("+mask" if relu_mask else "") + ("+drop" if drop is not None else ""), nbytes=_nbytes(x, dy, dx, w9c), flops=36 * x.numel())


# ------------------------------------------------------------------------------------------------ fused first block
def stem_supported(cin: int, cout: int) -> bool:
    return cin == 3 and cout == 64


def stem_fwd(x: torch.Tensor, wd9c: torch.Tensor, wp: torch.Tensor, out: torch.Tensor, scale=None, shift=None,
             relu: bool = False, colsum=None, colsq=None, d_out: Optional[torch.Tensor] = None) -> None:
    """enc1_block1_sepconv (u_net.py:63-66 on the RGB image): depthwise 3x3 + pointwise 3->64 in one kernel.
    d_out (fp32 [N,H,W,3]): also keep the depthwise output (input of stem_bwd_folded)."""
    n, h, w, cin, ldx = _nhwc(x, "x")
    _, _, _, cout, ldo = _nhwc(out, "out")
    if ldx != cin or x.dtype != out.dtype:
        raise ValueError("stem_fwd: x must be contiguous and share out's dtype")
    _f32(wd9c, "wd9c"); _f32(wp, "wp"); _f32(scale, "scale"); _f32(shift, "shift"); _f32(d_out, "d_out")
    if d_out is not None and d_out.numel() != n * h * w * cin:
        raise ValueError("stem_fwd: d_out must hold N*H*W*Cin floats")
    _call("unet_stem_fwd", _p(x), _p(wd9c), _p(wp), _p(out), ldo, n, h, w, cin, cout, _dt(x), _p(scale), _p(shift),
          int(relu), _p(colsum), _p(colsq), _p(d_out), _stream(), tag=f"{n}x{h}x{w}x{cin}->{cout}", nbytes=_nbytes(x, out, d_out),
          flops=(18 * cin + 2 * cin * cout) * n * h * w)


def stem_bwd(x: torch.Tensor, dz: torch.Tensor, wd9c, wp, dwd9c, dwp) -> None:
    n, h, w, cin, ldx = _nhwc(x, "x")
    _, _, _, cout, lddz = _nhwc(dz, "dz")
    if ldx != cin or x.dtype != dz.dtype:
        raise ValueError("stem_bwd: x must be contiguous and share dz's dtype")
    for t, nm in ((wd9c, "wd9c"), (wp, "wp"), (dwd9c, "dwd9c"), (dwp, "dwp")):
        _f32(t, nm)
    _call("unet_stem_bwd", _p(x), _p(dz), lddz, _p(wd9c), _p(wp), _p(dwd9c), _p(dwp), n, h, w, cin, cout, _dt(x),
          _stream(), tag=f"{n}x{h}x{w}x{cin}->{cout}", nbytes=_nbytes(x, dz), flops=(36 * cin + 4 * cin * cout) * n * h * w)


def stem_bwd_folded(g: torch.Tensor, z: torch.Tensor, coef: torch.Tensor, d3: torch.Tensor, wp, dwp, dd: torch.Tensor) -> None:
    """First block backward with BatchNormalization backward folded in: dwp += d3^T dz, dd = dz Wp^T, dz = A*g + B*z + K."""
    n, h, w, cout, ldg = _nhwc(g, "g")
    if not z.is_contiguous() or z.shape != g.shape or z.dtype != g.dtype or dd.dtype != g.dtype or not dd.is_contiguous():
        raise ValueError("stem_bwd_folded: z must be contiguous like g; dd contiguous of the same dtype")
    for t, nm in ((coef, "coef"), (d3, "d3"), (wp, "wp"), (dwp, "dwp")):
        _f32(t, nm)
    m = n * h * w
    if coef.numel() != 3 * cout or d3.numel() != 3 * m or dd.numel() != 3 * m:
        raise ValueError("stem_bwd_folded: coef [3,64], d3 [M,3], dd [M,3]")
    _call("unet_stem_bwd_folded", _p(g), ldg, _p(z), _p(coef), _p(d3), _p(wp), _p(dwp), _p(dd), m, _dt(g), _stream(),
          tag=f"{n}x{h}x{w}x3->{cout}", nbytes=_nbytes(g, z, d3, dd), flops=12 * 3 * cout * m // 3)


# ------------------------------------------------------------------------------------------------ fused conv_block (inference)
def sepconv_fused_supported(x: torch.Tensor, cout: int) -> bool:
    cin = x.shape[-1]
    return x.dtype == torch.bfloat16 and cin % 8 == 0 and cout % 8 == 0 and cin <= 256 and cout <= 128


def sepconv_fused(x: torch.Tensor, wd9c: torch.Tensor, wp_t: torch.Tensor, y: Optional[torch.Tensor],
                  scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None, relu: bool = True,
                  head_w: Optional[torch.Tensor] = None, head_b: Optional[torch.Tensor] = None,
                  head_out: Optional[torch.Tensor] = None, pooled: Optional[torch.Tensor] = None) -> None:
    """conv_block for inference in one kernel: depthwise 3x3 produced on chip as the tcgen05 A operand, pointwise GEMM,
    folded BN + ReLU epilogue, TMA store into the (possibly channel-sliced) destination; optionally the 1x1 output head
    (sigmoid / softmax) from the same registers, in which case `y` may be None."""
    n, h, w, cin, ldx = _nhwc(x, "x")
    cout = wp_t.shape[0]
    ldy = 0
    if y is not None:
        n2, h2, w2, c2, ldy = _nhwc(y, "y")
        if (n, h, w, cout) != (n2, h2, w2, c2) or y.dtype != torch.bfloat16:
            raise ValueError("sepconv_fused: x / y disagree or y is not bf16")
    if x.dtype != torch.bfloat16 or wp_t.dtype != torch.bfloat16 or wp_t.shape[1] != cin or wp_t.stride(1) != 1:
        raise ValueError("sepconv_fused: x must be bf16 and wp_t bf16 [Cout, Cin]")
    _f32(wd9c, "wd9c"); _f32(scale, "scale"); _f32(shift, "shift"); _f32(head_w, "head_w"); _f32(head_b, "head_b"); _f32(head_out, "head_out")
    classes = 0
    if head_out is not None:
        if head_w is None or head_w.shape[0] != cout or head_out.numel() != n * h * w * head_w.shape[1]:
            raise ValueError("sepconv_fused: head_w must be [Cout, classes] and head_out [N,H,W,classes]")
        classes = head_w.shape[1]
    elif y is None:
        raise ValueError("sepconv_fused: nothing to produce")
    ldp = 0
    if pooled is not None:     # MaxPooling2D((2,2)) of the stored activation from the same staged tile
        n3, h3, w3, c3, ldp = _nhwc(pooled, "pooled")
        if y is None or (n3, h3 * 2, w3 * 2, c3) != (n, h, w, cout) or pooled.dtype != torch.bfloat16:
            raise ValueError("sepconv_fused: pooled must be a bf16 [N,H/2,W/2,Cout] view and needs y")
    _call("unet_sepconv_fused_fwd", _p(x), ldx, _p(wd9c), _p(wp_t), wp_t.stride(0), _p(scale), _p(shift), int(relu), _p(y), ldy,
          n, h, w, cin, cout, _p(head_w), _p(head_b), _p(head_out), classes, _p(pooled), ldp, _stream(),
          tag=f"{n}x{h}x{w}x{cin}->{cout}{'+head' if head_out is not None else ''}{'+pool' if pooled is not None else ''}",
          nbytes=_nbytes(x, y, wp_t, head_out, pooled), flops=(18 * cin + 2 * cin * cout) * n * h * w)


# ------------------------------------------------------------------------------------------------ dense contractions
_split_ws = {}    # device index -> flat fp32 workspace for the (hi, lo) split of the A operand (fp32 mode on the tensor cores)


def split_tf32(src: torch.Tensor, hi: Optional[torch.Tensor], lo: torch.Tensor, transpose: bool = False) -> None:
    """lo = tf32(src - trunc13(src)): the part of src the tensor cores drop when they read it as tf32 (src itself is the `hi`
    operand).  hi (optional) receives src unchanged; hi / lo contiguous [rows, cols] ([cols, rows] with transpose)."""
    rows, cols, ld = _rows(src, "src")
    _f32(hi, "hi"); _f32(lo, "lo")
    if src.dtype != torch.float32 or lo.numel() != rows * cols or (hi is not None and hi.numel() != rows * cols):
        raise ValueError("split_tf32: src must be fp32 and hi / lo hold rows*cols floats")
    _call("unet_split_tf32", _p(src), ld, rows, cols, _p(hi), _p(lo), int(transpose), _stream(),
          tag=f"{rows}x{cols}", nbytes=(2 if hi is None else 3) * rows * cols * 4)


def _a_split(A: torch.Tensor, rows: int, cols: int, slot: int = 0) -> torch.Tensor:
    key = (A.device.index or 0, slot)
    ws = _split_ws.get(key)
    if ws is None or ws.numel() < rows * cols:
        ws = _split_ws[key] = torch.empty(rows * cols, device=A.device, dtype=torch.float32)
    lo = ws[: rows * cols].view(rows, cols)
    split_tf32(A, None, lo)
    return lo


def tf32_lo(t: torch.Tensor, slot: int) -> torch.Tensor:
    """The tf32 `lo` part of an fp32 operand in workspace `slot` (contiguous [rows, cols]), for handing the same split to
    several gemm calls (`A_lo=` / `B_lo=`): a gradient tensor feeds both its weight-gradient and its data-gradient GEMM."""
    rows, cols, _ = _rows(t, "t")
    return _a_split(t, rows, cols, slot=slot)


def gemm(A: torch.Tensor, B: torch.Tensor, Cm: Optional[torch.Tensor], *, a_trans: bool = False, b_trans: bool = False,
         accumulate: bool = False, epilogue: int = EPI_NONE, scale: Optional[torch.Tensor] = None,
         shift: Optional[torch.Tensor] = None, colsum: Optional[torch.Tensor] = None,
         colsq: Optional[torch.Tensor] = None, convt_hw: Tuple[int, int] = (0, 0),
         drop: Optional[Dropout] = None, tensor_core: Optional[bool] = None,
         head_w: Optional[torch.Tensor] = None, head_b: Optional[torch.Tensor] = None,
         head_out: Optional[torch.Tensor] = None, A2: Optional[torch.Tensor] = None,
         B2: Optional[torch.Tensor] = None, B_lo: Optional[torch.Tensor] = None, tf32x3: bool = False,
         A_lo: Optional[torch.Tensor] = None) -> None:
    """C[M,N] (+)= op(A) op(B) with a fused epilogue.  bf16 operands go to the tcgen05 kernel when its layout rules
    hold (forward/dgrad: B given as [N,K]; weight gradient: a_trans, accumulate), everything else to the fp32-exact
    CUDA-core kernel.  `tensor_core` forces the choice (True raises if the layout is not supported).
    A2 (a_trans=False): the A operand is [A | A2] along K;  B2 (a_trans=True): the B operand is [B | B2] along N
    (tensor-core path only; the first part must be a multiple of 64 columns wide).
    B_lo: fp32 operands on the tensor cores — B_lo is the `lo` part of the [N,K] operand B (split_tf32); A's lo part is
    written into a workspace here; three kind::tf32 MMAs per k-step give fp32-grade products.
    tf32x3 (a_trans=True, accumulate): the fp32 weight gradient on the tensor cores; both operands are split here."""
    ar, ac, lda = _rows(A, "A")
    br, bc, ldb = _rows(B, "B")
    M, K = (ac, ar) if a_trans else (ar, ac)
    N, Kb = (br, bc) if b_trans else (bc, br)
    lda2 = ldb2 = k_split = n_split = 0
    if A2 is not None:
        if a_trans or A2.dtype != A.dtype:
            raise ValueError("gemm: A2 needs a_trans=False and the dtype of A")
        a2r, a2c, lda2 = _rows(A2, "A2")
        if a2r != M:
            raise ValueError("gemm: A and A2 must have the same number of rows")
        k_split, K = K, K + a2c
    if B2 is not None:
        if not a_trans or b_trans or B2.dtype != B.dtype:
            raise ValueError("gemm: B2 needs a_trans=True, b_trans=False and the dtype of B")
        b2r, b2c, ldb2 = _rows(B2, "B2")
        if b2r != Kb:
            raise ValueError("gemm: B and B2 must have the same number of rows")
        n_split, N = N, N + b2c
    if K != Kb:
        raise ValueError(f"gemm: inner dimensions disagree ({K} vs {Kb})")
    if A.dtype != B.dtype:
        raise ValueError("gemm: A and B must share a dtype")
    if Cm is None:
        if epilogue != EPI_HEAD:
            raise ValueError("gemm: C may be omitted only with the fused output head")
        ldc = N
    elif epilogue == EPI_CONVT:
        _, _, _, cc, ldc = _nhwc(Cm, "C")
        if cc * 4 != N:
            raise ValueError("gemm(CONVT): destination view must have N/4 channels")
    else:
        cr, cc, ldc = _rows(Cm, "C")
        if (cr, cc) != (M, N):
            raise ValueError(f"gemm: C is {cr}x{cc}, expected {M}x{N}")
    _f32(scale, "scale"); _f32(shift, "shift")
    for t, nm in ((colsum, "colsum"), (colsq, "colsq")):
        if t is not None and t.dtype != torch.float64:
            raise TypeError(f"{nm} must be float64")
    args = GemmArgs()
    args.M, args.N, args.K = M, N, K
    args.A, args.lda = _p(A), lda
    args.B, args.ldb = _p(B), ldb
    args.C, args.ldc = _p(Cm), ldc
    args.a_trans, args.b_trans = int(a_trans), int(b_trans)
    args.in_dtype, args.out_dtype = _dt(A), (_dt(Cm) if Cm is not None else _dt(A))
    args.accumulate, args.epilogue = int(accumulate), int(epilogue)
    args.scale, args.shift = _p(scale), _p(shift)
    args.colsum, args.colsq = _p(colsum), _p(colsq)
    args.convt_H, args.convt_W = int(convt_hw[0]), int(convt_hw[1])
    if drop is not None:
        args.drop = drop
    if epilogue == EPI_HEAD:
        _f32(head_w, "head_w"); _f32(head_b, "head_b"); _f32(head_out, "head_out")
        if head_w is None or head_out is None or head_w.shape[0] != N or head_out.numel() != M * head_w.shape[1]:
            raise ValueError("gemm(HEAD): head_w must be [N, classes] and head_out [M, classes]")
        args.head_w, args.head_b, args.head_out, args.head_classes = _p(head_w), _p(head_b), _p(head_out), head_w.shape[1]
    args.A2, args.lda2, args.k_split = _p(A2), lda2, k_split
    args.B2, args.ldb2, args.n_split = _p(B2), ldb2, n_split
    tc_ok = (A.dtype == torch.bfloat16 and N % 8 == 0 and lda % 8 == 0 and ldb % 8 == 0 and ldc % 4 == 0
             and ((not a_trans and b_trans and not accumulate and K % 8 == 0)
                  or (a_trans and not b_trans and accumulate and M % 8 == 0))
             and (epilogue != EPI_CONVT or ((N // 4) % 64 == 0 and convt_hw[1] > 0 and
                                           (128 % convt_hw[1] == 0 or convt_hw[1] % 128 == 0))))
    use_tc = tc_ok if tensor_core is None else tensor_core
    if (A.dtype == torch.float32 and B_lo is not None and tensor_core is not False and not a_trans and b_trans and not accumulate
            and A2 is None and B2 is None and epilogue != EPI_HEAD and Cm is not None and Cm.dtype == torch.float32
            and K % 4 == 0 and N % 8 == 0 and lda % 4 == 0 and ldb % 4 == 0 and ldc % 4 == 0
            and (epilogue != EPI_CONVT or ((N // 4) % 32 == 0 and convt_hw[1] > 0 and
                                           (128 % convt_hw[1] == 0 or convt_hw[1] % 128 == 0)
                                           and (drop is None or (drop.ctot % 4 == 0 and drop.c0 % 4 == 0))))):
        if B_lo.dtype != torch.float32 or tuple(B_lo.shape) != tuple(B.shape) or B_lo.stride() != B.stride():
            raise ValueError("gemm: B_lo must match B (fp32, same shape and strides)")
        lo = A_lo if A_lo is not None else _a_split(A, M, K)
        if tuple(lo.shape) != (M, K) or not lo.is_contiguous() or lo.dtype != torch.float32:
            raise ValueError("gemm: A_lo must be a contiguous fp32 [M, K] tensor")
        args.A_lo, args.lda_lo, args.B_lo, args.ldb_lo = _p(lo), K, _p(B_lo), ldb
        csz = Cm.numel() * 4
        _call("unet_gemm_tc", C.byref(args), _stream(),
              tag=f"{'convt' if epilogue == EPI_CONVT else 'nt'}:{M}x{N}x{K}:e{epilogue}:tf32x3",
              nbytes=_nbytes(A, B) + csz, flops=2 * M * N * K)
        return
    if (tf32x3 and A.dtype == torch.float32 and tensor_core is not False and a_trans and not b_trans and accumulate
            and A2 is None and B2 is None and epilogue == EPI_NONE and Cm is not None and Cm.dtype == torch.float32
            and M % 4 == 0 and N % 8 == 0 and lda % 4 == 0 and ldb % 4 == 0):
        a_lo = A_lo if A_lo is not None else _a_split(A, K, M, slot=0)
        b_lo = B_lo if B_lo is not None else _a_split(B, K, N, slot=1)
        if tuple(a_lo.shape) != (K, M) or tuple(b_lo.shape) != (K, N) or not a_lo.is_contiguous() or not b_lo.is_contiguous():
            raise ValueError("gemm: A_lo / B_lo must be contiguous [K, M] / [K, N] tensors")
        args.A_lo, args.lda_lo, args.B_lo, args.ldb_lo = _p(a_lo), M, _p(b_lo), N
        _call("unet_gemm_tc", C.byref(args), _stream(), tag=f"wgrad:{M}x{N}x{K}:e0:tf32x3",
              nbytes=_nbytes(A, B) + 2 * Cm.numel() * 4, flops=2 * M * N * K)
        return
    if A.dtype == torch.float32 and tensor_core is True:
        raise ValueError("gemm: fp32 operands reach the tensor cores only as C = A * B^T with B_lo given (tf32 split) and aligned shapes")
    csz = (Cm.numel() * Cm.element_size() if Cm is not None else 0) + (head_out.numel() * 4 if head_out is not None else 0)
    _call("unet_gemm_tc" if use_tc else "unet_gemm_simt", C.byref(args), _stream(),
          tag=f"{'wgrad' if a_trans else ('convt' if epilogue == EPI_CONVT else 'nt')}:{M}x{N}x{K}:e{epilogue}",
          nbytes=_nbytes(A, B, A2, B2) + csz * (2 if accumulate else 1),
          flops=2 * M * N * K)


def pw_bwd_fused_supported(g, z, d, dd) -> bool:
    """unet_pw_bwd_fused takes bf16 row views with C == 64 and Cin in {64, 128}."""
    if any(t.dtype != torch.bfloat16 for t in (g, z, d, dd)):
        return False
    c, cin = g.shape[-1], d.shape[-1]
    return c == 64 and cin in (64, 128) and z.shape[-1] == c and dd.shape[-1] == cin


def pw_bwd_fused(g, z, d, wab, bias, dd, G) -> None:
    """Folded pointwise backward in one pass over [g | z] and d:  dd = [g | z] wab^T + bias  and  G += d^T [g | z]
    (g, z: [P,C]; d, dd: [P,Cin]; wab bf16 [Cin,2C]; bias fp32 [Cin]; G fp32 [Cin,2C], accumulated)."""
    P, c, ldg = _rows(g, "g")
    pz, cz, ldz = _rows(z, "z")
    pd, cin, ldd = _rows(d, "d")
    po, co, ldo = _rows(dd, "dd")
    if not (P == pz == pd == po) or cz != c or co != cin:
        raise ValueError("pw_bwd_fused: g, z [P,C] and d, dd [P,Cin] must agree")
    if wab.dtype != torch.bfloat16 or tuple(wab.shape) != (cin, 2 * c) or not wab.is_contiguous():
        raise ValueError("pw_bwd_fused: wab must be a contiguous bf16 [Cin, 2C]")
    _f32(bias, "bias"); _f32(G, "G")
    if tuple(G.shape) != (cin, 2 * c) or not G.is_contiguous() or bias.numel() != cin:
        raise ValueError("pw_bwd_fused: G must be a contiguous fp32 [Cin, 2C] and bias [Cin]")
    _call("unet_pw_bwd_fused", _p(g), ldg, _p(z), ldz, _p(d), ldd, _p(wab), 2 * c, _p(bias), _p(dd), ldo, _p(G), 2 * c,
          P, cin, c, _stream(), tag=f"{P}x{cin}x{2 * c}", nbytes=_nbytes(g, z, d, dd), flops=2 * 2 * P * cin * 2 * c)


# ------------------------------------------------------------------------------------------------ batch normalisation
def bn_bwd_coef(sums, gamma, beta, save_mean, save_rstd, count: int, dgamma, dbeta, coef=None, w=None, wab=None, bias=None,
                guard: Optional[torch.Tensor] = None) -> None:
    """BatchNormalization backward as per-channel coefficients dz = A*g + B*z + K (coef fp32 [3,C]); accumulates dgamma/dbeta;
    with the pointwise kernel w [Cin,C] also the folded data-gradient operands wab (bf16 [Cin,2C]) and bias [Cin].
    guard (CUDA int32 [1], optional): set to 1 by the kernel when the fold is ill-conditioned (|gamma| < |beta|/16)."""
    if guard is not None and (guard.dtype != torch.int32 or not guard.is_cuda):
        raise TypeError("guard must be a CUDA int32 tensor")
    for t, nm in ((sums, "sums"), (gamma, "gamma"), (beta, "beta"), (save_mean, "save_mean"), (save_rstd, "save_rstd"),
                  (dgamma, "dgamma"), (dbeta, "dbeta"), (coef, "coef"), (w, "w"), (bias, "bias")):
        _f32(t, nm)
    c = gamma.numel()
    cin = 0
    if w is not None:
        cin = w.shape[0]
        if tuple(w.shape) != (cin, c) or wab is None or wab.dtype != torch.bfloat16 or tuple(wab.shape) != (cin, 2 * c) \
                or not wab.is_contiguous() or bias is None or bias.numel() != cin:
            raise ValueError("bn_bwd_coef: w [Cin,C] needs wab bf16 [Cin,2C] and bias [Cin]")
    if sums.numel() != 2 * c or (coef is not None and coef.numel() != 3 * c):
        raise ValueError("bn_bwd_coef: sums must hold 2*C and coef 3*C floats")
    _call("unet_bn_bwd_coef", _p(sums), _p(gamma), _p(beta), _p(save_mean), _p(save_rstd), int(count), _p(dgamma), _p(dbeta),
          _p(coef), _p(w), cin, c, _p(wab), _p(bias), _p(guard), _stream())


def bn_bwd_wgrad_combine(G, coef, sd, dw) -> None:
    """dw[i,c] += G[i,c]*A[c] + G[i,C+c]*B[c] + sd[i]*K[c]  (G = d^T [g | z] fp32 [Cin,2C])."""
    for t, nm in ((G, "G"), (coef, "coef"), (sd, "sd"), (dw, "dw")):
        _f32(t, nm)
    cin, c = dw.shape
    if tuple(G.shape) != (cin, 2 * c) or coef.numel() != 3 * c or sd.numel() != cin:
        raise ValueError("bn_bwd_wgrad_combine: shapes disagree")
    _call("unet_bn_bwd_wgrad_combine", _p(G), _p(coef), _p(sd), _p(dw), cin, c, _stream())


def bn_fold(gamma, beta, mean, var, eps: float, scale, shift) -> None:
    for t, nm in ((gamma, "gamma"), (beta, "beta"), (mean, "mean"), (var, "var"), (scale, "scale"), (shift, "shift")):
        _f32(t, nm)
    _call("unet_bn_fold", _p(gamma), _p(beta), _p(mean), _p(var), float(eps), _p(scale), _p(shift),
          scale.numel(), _stream())


def bn_finalize(colsum, colsq, count: int, gamma, beta, eps: float, momentum: float, moving_mean, moving_var,
                scale, shift, save_mean, save_rstd) -> None:
    for t, nm in ((gamma, "gamma"), (beta, "beta"), (moving_mean, "moving_mean"), (moving_var, "moving_var"),
                  (scale, "scale"), (shift, "shift"), (save_mean, "save_mean"), (save_rstd, "save_rstd")):
        _f32(t, nm)
    _call("unet_bn_finalize", _p(colsum), _p(colsq), int(count), _p(gamma), _p(beta), float(eps), float(momentum),
          _p(moving_mean), _p(moving_var), _p(scale), _p(shift), _p(save_mean), _p(save_rstd), scale.numel(), _stream())


def bn_act(z: torch.Tensor, scale, shift, y: torch.Tensor, relu: bool = True, pooled: Optional[torch.Tensor] = None,
           drop: Optional[Dropout] = None) -> None:
    n, h, w, c, ldz = _nhwc(z, "z")
    if ldz != c:
        raise ValueError("bn_act: z must be contiguous")
    _, _, _, c2, ldy = _nhwc(y, "y")
    if c2 != c or (pooled is not None and not pooled.is_contiguous()):
        raise ValueError("bn_act: bad y / pooled")
    _f32(scale, "scale"); _f32(shift, "shift")
    _call("unet_bn_act", _p(z), _p(scale), _p(shift), int(relu), _p(y), ldy, _p(pooled), n, h, w, c, _dt(z),
          _dref(drop), _stream(), tag=f"{n}x{h}x{w}x{c}{'+pool' if pooled is not None else ''}",
          nbytes=_nbytes(z, y, pooled))


def bn_bwd_reduce(dy, z, scale, shift, save_mean, save_rstd, dgamma, dbeta, relu: bool = True,
                  drop: Optional[Dropout] = None) -> None:
    M, c, lddy = _rows(dy, "dy")
    M2, c2, ldz = _rows(z, "z")
    if (M, c) != (M2, c2) or ldz != c:
        raise ValueError("bn_bwd_reduce: dy / z disagree or z not contiguous")
    _call("unet_bn_bwd_reduce", _p(dy), lddy, _p(z), _p(scale), _p(shift), _p(save_mean), _p(save_rstd),
          _p(dgamma), _p(dbeta), M, c, _dt(z), int(relu), _dref(drop), _stream(), tag=f"{M}x{c}", nbytes=_nbytes(dy, z))


def bn_bwd_apply(dy, z, scale, shift, save_mean, save_rstd, dgamma, dbeta, dz, relu: bool = True,
                 drop: Optional[Dropout] = None) -> None:
    M, c, lddy = _rows(dy, "dy")
    M2, c2, ldz = _rows(z, "z")
    if (M, c) != (M2, c2) or ldz != c or not dz.is_contiguous():
        raise ValueError("bn_bwd_apply: dy / z disagree or z, dz not contiguous")
    _call("unet_bn_bwd_apply", _p(dy), lddy, _p(z), _p(scale), _p(shift), _p(save_mean), _p(save_rstd),
          _p(dgamma), _p(dbeta), _p(dz), M, c, _dt(z), int(relu), _dref(drop), _stream(), tag=f"{M}x{c}",
          nbytes=_nbytes(dy, z, dz))


# ------------------------------------------------------------------------------------------------ pooling
def maxpool2x2(x: torch.Tensor, y: torch.Tensor) -> None:
    n, h, w, c, ldx = _nhwc(x, "x")
    if tuple(y.shape) != (n, h // 2, w // 2, c) or not y.is_contiguous():
        raise ValueError("maxpool2x2: y must be contiguous (N,H/2,W/2,C)")
    _call("unet_maxpool2x2_fwd", _p(x), ldx, _p(y), n, h, w, c, _dt(x), _stream(), tag=f"{n}x{h}x{w}x{c}",
          nbytes=_nbytes(x, y))


def maxpool2x2_bwd(z: torch.Tensor, scale, shift, dpool: torch.Tensor, dskip: Optional[torch.Tensor],
                   dy: torch.Tensor, bn_sums: Optional[torch.Tensor] = None, skip_drop: Optional[Dropout] = None) -> None:
    n, h, w, c, ldz = _nhwc(z, "z")
    lds = 0
    if dskip is not None:
        lds = _nhwc(dskip, "dskip")[4]
    if not dpool.is_contiguous() or not dy.is_contiguous():
        raise ValueError("maxpool2x2_bwd: dpool and dy must be contiguous")
    _f32(bn_sums, "bn_sums")
    if bn_sums is not None and bn_sums.numel() != 2 * c:
        raise ValueError("maxpool2x2_bwd: bn_sums must hold 2*C floats")
    _call("unet_maxpool2x2_bwd", _p(z), ldz, _p(scale), _p(shift), _p(dpool), _p(dskip), lds, _p(dy), n, h, w, c,
          _dt(z), _p(bn_sums), _dref(skip_drop), _stream(), tag=f"{n}x{h}x{w}x{c}", nbytes=_nbytes(z, dpool, dskip, dy))


def convt_bwd_gather(du: torch.Tensor, g: torch.Tensor, dbias: Optional[torch.Tensor], drop: Optional[Dropout] = None) -> None:
    """du: (N,2H,2W,Cout) view -> g: [N*H*W, 4*Cout]; dbias += column sums."""
    n, h2, w2, co, lddu = _nhwc(du, "du")
    if not g.is_contiguous() or g.numel() != du.shape[0] * h2 * w2 * co:
        raise ValueError("convt_bwd_gather: g has the wrong size")
    _call("unet_convt_bwd_gather", _p(du), lddu, _p(g), _p(dbias), n, h2 // 2, w2 // 2, co, _dt(du), _dref(drop), _stream(),
          tag=f"{n}x{h2}x{w2}x{co}", nbytes=_nbytes(du, g))


# ------------------------------------------------------------------------------------------------ head + loss
def head_fwd(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], probs: torch.Tensor,
             y_true: Optional[torch.Tensor] = None, sums: Optional[torch.Tensor] = None,
             x_scale: Optional[torch.Tensor] = None, x_shift: Optional[torch.Tensor] = None) -> None:
    n, h, wd, k, ldx = _nhwc(x, "x")
    c = probs.shape[-1]
    _f32(w, "w"); _f32(b, "b"); _f32(probs, "probs"); _f32(y_true, "y_true"); _f32(x_scale, "x_scale"); _f32(x_shift, "x_shift")
    if sums is not None and sums.dtype != torch.float64:
        raise TypeError("sums must be float64")
    _call("unet_head_fwd", _p(x), ldx, _p(w), _p(b), _p(probs), _p(y_true), _p(sums), n * h * wd, h * wd, k, c,
          _dt(x), _p(x_scale), _p(x_shift), _stream(), tag=f"{n}x{h}x{wd}x{k}->{c}", nbytes=_nbytes(x, probs, y_true), flops=2 * x.numel() * c)


def seg_loss_finalize(sums: torch.Tensor, npairs: int, smooth: float, kind: int, grad_scale: float,
                      out3: torch.Tensor, coef: Optional[torch.Tensor]) -> None:
    _f32(out3, "out3"); _f32(coef, "coef")
    _call("unet_seg_loss_finalize", _p(sums), int(npairs), float(smooth), int(kind), float(grad_scale), _p(out3),
          _p(coef), _stream())


def head_stream_supported(x: torch.Tensor, classes: int) -> bool:
    """the streamed binary-head kernels (and with them BN+ReLU on load) apply"""
    return x.dtype == torch.bfloat16 and classes == 1 and x.shape[-1] == 64 and x.is_contiguous()


def head_bwd(x, w, probs, y_true, coef, dx: Optional[torch.Tensor], dw, db, bn_sums: Optional[torch.Tensor] = None,
             x_scale: Optional[torch.Tensor] = None, x_shift: Optional[torch.Tensor] = None) -> None:
    n, h, wd, k, ldx = _nhwc(x, "x")
    c = probs.shape[-1]
    lddx = _nhwc(dx, "dx")[4] if dx is not None else 0
    for t, nm in ((w, "w"), (probs, "probs"), (y_true, "y_true"), (coef, "coef"), (dw, "dw"), (db, "db")):
        _f32(t, nm)
    _f32(bn_sums, "bn_sums"); _f32(x_scale, "x_scale"); _f32(x_shift, "x_shift")
    if bn_sums is not None and bn_sums.numel() != 2 * k:
        raise ValueError("head_bwd: bn_sums must hold 2*K floats")
    _call("unet_head_bwd", _p(x), ldx, _p(w), _p(probs), _p(y_true), _p(coef), _p(dx), lddx, _p(dw), _p(db),
          n * h * wd, h * wd, k, c, _dt(x), _p(bn_sums), _p(x_scale), _p(x_shift), _stream(), tag=f"{n}x{h}x{wd}x{k}->{c}",
          nbytes=_nbytes(x, probs, y_true, dx), flops=4 * x.numel() * c)


def seg_sums(y_true: torch.Tensor, y_pred: torch.Tensor, sums: torch.Tensor) -> None:
    _f32(y_true, "y_true"); _f32(y_pred, "y_pred")
    nb, c = y_true.shape[0], y_true.shape[-1]
    hw = y_true.numel() // (nb * c)
    _call("unet_seg_sums", _p(y_true), _p(y_pred), _p(sums), nb, hw, c, _stream())


def confusion_matrix_update(y_true, y_pred, num_classes: int, counts: torch.Tensor,
                            threshold: Optional[float] = None) -> None:
    _f32(y_true, "y_true"); _f32(y_pred, "y_pred")
    if counts.dtype != torch.int64 or not counts.is_contiguous():
        raise TypeError("counts must be contiguous int64")
    if threshold is None:
        _call("unet_confusion_matrix_update", _p(y_true), _p(y_pred), y_true.numel(), int(num_classes), _p(counts),
              _stream())
    else:
        _call("unet_confusion_matrix_update_thr", _p(y_true), _p(y_pred), float(threshold), y_true.numel(),
              _p(counts), _stream())


def sample_confusion_thr(y_true: torch.Tensor, prob: torch.Tensor, threshold: float, counts: torch.Tensor) -> None:
    """Per-sample 2x2 confusion counts of (prob > threshold) against a {0,1} truth: counts [NB,4] int64, accumulated
    (calculate_sample_iou + MeanIoU(2) of scripts/benchmark.py:159-170,260,269 from one pass over the batch)."""
    _f32(y_true, "y_true"); _f32(prob, "prob")
    nb = y_true.shape[0]
    if y_true.numel() != prob.numel() or prob.shape[0] != nb:
        raise ValueError("sample_confusion_thr: y_true and prob disagree")
    if counts.dtype != torch.int64 or not counts.is_contiguous() or tuple(counts.shape) != (nb, 4) or not counts.is_cuda:
        raise TypeError("counts must be a contiguous CUDA int64 [NB,4] tensor")
    _call("unet_sample_confusion_thr", _p(y_true), _p(prob), float(threshold), nb, y_true.numel() // nb, _p(counts), _stream())


# ------------------------------------------------------------------------------------------------ optimiser, staging
def adamw_step(w, g, m, v, hyper) -> None:
    for t, nm in ((w, "w"), (g, "g"), (m, "m"), (v, "v"), (hyper, "hyper")):
        _f32(t, nm)
    _call("unet_adamw_step", _p(w), _p(g), _p(m), _p(v), w.numel(), _p(hyper), _stream(), tag=str(w.numel()),
          nbytes=7 * w.numel() * 4)


def step_advance(hyper: Optional[torch.Tensor], counter: Optional[torch.Tensor]) -> None:
    """hyper[5] (AdamW step t) += 1; *counter (dropout seed word) += 1 — on device, so CUDA-graph replays advance."""
    _f32(hyper, "hyper")
    _call("unet_step_advance", _p(hyper), _p(counter), _stream())


def cast_transpose_bf16(src: torch.Tensor, dst: Optional[torch.Tensor], dst_t: Optional[torch.Tensor],
                        col_scale: Optional[torch.Tensor] = None) -> None:
    """dst[r,c] = bf16(src[r,c] * col_scale[c]), dst_t = its transpose (either may be None)."""
    _f32(src, "src"); _f32(col_scale, "col_scale")
    r, c = src.shape
    if col_scale is not None and col_scale.numel() != c:
        raise ValueError("cast_transpose_bf16: col_scale must hold one value per column")
    _call("unet_cast_transpose_bf16", _p(src), _p(dst), _p(dst_t), r, c, _p(col_scale), _stream())


def cast_transpose_table(base: torch.Tensor, items) -> tuple:
    """Device table for cast_transpose_bf16_batched.  items: [(offset of the [R, C] fp32 matrix in `base` (elements), dst or None,
    dst_t or None, R, C)]; destinations are contiguous bf16 [R, C] / [C, R] tensors the caller keeps alive.
    Returns (table, n, total_tiles)."""
    rows, tile0 = [], 0
    for off, dst, dst_t, r, c in items:
        for t, shape in ((dst, (r, c)), (dst_t, (c, r))):
            if t is not None and (t.dtype != torch.bfloat16 or tuple(t.shape) != shape or not t.is_contiguous()):
                raise ValueError("cast_transpose_table: destinations must be contiguous bf16 [R, C] / [C, R]")
        if off < 0 or off + r * c > base.numel():
            raise ValueError("cast_transpose_table: matrix outside the source buffer")
        rows.append([off, _p(dst) or 0, _p(dst_t) or 0, r, c, tile0])
        tile0 += ((r + 31) // 32) * ((c + 31) // 32)
    table = torch.tensor(rows, dtype=torch.int64, device=base.device)
    return table, len(rows), tile0


def cast_transpose_bf16_batched(base: torch.Tensor, table: torch.Tensor, n: int, total_tiles: int) -> None:
    """Every matrix of a cast_transpose_table in one launch: dst = bf16(src), dst_t = bf16(src)^T."""
    _f32(base, "base")
    if table.dtype != torch.int64 or not table.is_contiguous() or table.numel() != 6 * n:
        raise ValueError("cast_transpose_bf16_batched: table must be int64 [n, 6]")
    _call("unet_cast_transpose_bf16_batched", _p(base), _p(table), n, total_tiles, _stream())


def cast(src: torch.Tensor, dst: torch.Tensor) -> None:
    if not src.is_contiguous() or not dst.is_contiguous() or src.numel() != dst.numel():
        raise ValueError("cast: tensors must be contiguous and equal in size")
    _call("unet_cast", _p(src), _dt(src), _p(dst), _dt(dst), src.numel(), _stream(), tag=str(src.numel()),
          nbytes=_nbytes(src, dst))


def device_check(device: int = 0) -> None:
    _lib.call("unet_device_check", int(device))


# ------------------------------------------------------------------------------------------------ CLI pre/post-processing
def preprocess_u8(img: torch.Tensor, out: torch.Tensor, divisor: float = 255.0) -> None:
    """img: CUDA uint8 [H0,W0,C] (cv2.imread order) -> out: fp32 [h,w,C] = cv2.resize(img/255, INTER_LINEAR)."""
    if img.dtype != torch.uint8 or img.dim() != 3 or not img.is_cuda or img.stride(2) != 1 or img.stride(1) != img.shape[2]:
        raise TypeError("preprocess_u8: img must be a CUDA uint8 [H,W,C] tensor with packed pixels")
    _f32(out, "out")
    h, w, c = out.shape[-3:]
    if c != img.shape[2]:
        raise ValueError("preprocess_u8: channel mismatch")
    _call("unet_preprocess_u8", _p(img), img.shape[0], img.shape[1], c, img.stride(0), _p(out), h, w, float(divisor), _stream())


def postprocess_mask(prob: torch.Tensor, mask: torch.Tensor, threshold: float) -> None:
    """prob: fp32 view [h,w] (any element stride) -> mask: CUDA uint8 [H0,W0] = (cv2.resize(prob) > threshold) * 255."""
    if prob.dtype != torch.float32 or prob.dim() != 2 or not prob.is_cuda:
        raise TypeError("postprocess_mask: prob must be a CUDA fp32 [h,w] view")
    ld = prob.stride(1) if prob.shape[1] > 1 else 1
    if prob.shape[0] > 1 and prob.stride(0) != prob.shape[1] * ld:
        raise ValueError("postprocess_mask: rows must be densely packed at the element stride")
    if mask.dtype != torch.uint8 or not mask.is_contiguous() or not mask.is_cuda:
        raise TypeError("postprocess_mask: mask must be a contiguous CUDA uint8 tensor")
    _call("unet_postprocess_mask", _p(prob), prob.shape[0], prob.shape[1], ld, _p(mask), mask.shape[0], mask.shape[1],
          float(threshold), _stream())
