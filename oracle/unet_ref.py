"""CPU oracle (TEST INFRASTRUCTURE ONLY) — NumPy restatement of the reference hot path.

PARITY UNPINNED: the reference (planck-epoch/unet-image-segmentation) has no tests, no golden tensors and no
shipped weights, and its arithmetic lives in TensorFlow/Keras, which is neither vendored nor pinned
(requirements.txt:15-19) and is not installable in this image.  This file restates (a) the reference's own code
line by line where it is plain tensor algebra (utils/metrics.py, utils/loss.py, model/u_net.py topology) and
(b) the published Keras/TF semantics of the layers the reference instantiates.  It is
cross-checked against an independent torch-CPU restatement with autograd (oracle/torch_ref.py) in tests/.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this package.
The product (unet-image-segmentation_b200/, model/, utils/, scripts/) never does.

Everything is NHWC.  `dtype` is np.float64 (truth) or np.float32 (what TF-CPU computes, up to summation order).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import numpy as np

EPSILON = 1e-7          # K.epsilon(), utils/metrics.py:4, utils/loss.py:7
BN_EPS = 1e-3           # keras BatchNormalization default epsilon
BN_MOMENTUM = 0.99      # keras BatchNormalization default momentum
FILTERS = [64, 128, 256, 512]   # model/u_net.py:57


# ----------------------------------------------------------------------------------------------- topology
def layer_specs(input_size: Tuple[int, int, int], num_classes: int = 1, dropout_rate: float = 0.2,
                use_batch_norm: bool = True) -> List[dict]:
    """Keras layer list in creation order, following model/u_net.py:28-116 (names are API)."""
    if len(input_size) != 3:
        raise ValueError("input_size must be a tuple of (height, width, channels)")   # u_net.py:52-53
    h, w, cin = input_size
    specs: List[dict] = [dict(kind="input", name="input_image", shape=(h, w, cin))]

    def conv_block(prefix, ci, co):      # u_net.py:5-26
        specs.append(dict(kind="sepconv", name=f"{prefix}_sepconv", cin=ci, cout=co, use_bias=not use_batch_norm))
        if use_batch_norm:
            specs.append(dict(kind="bn", name=f"{prefix}_bn", c=co))
        specs.append(dict(kind="relu", name=f"{prefix}_relu"))

    c = cin
    for i, f in enumerate(FILTERS):      # encoder, u_net.py:63-69
        s = i + 1
        conv_block(f"enc{s}_block1", c, f)
        conv_block(f"enc{s}_block2", f, f)
        specs.append(dict(kind="pool", name=f"enc{s}_pool"))
        c = f
    bn_f = FILTERS[-1] * 2               # bottleneck, u_net.py:73-78
    conv_block("bneck_block1", c, bn_f)
    conv_block("bneck_block2", bn_f, bn_f)
    if dropout_rate > 0.0:
        specs.append(dict(kind="dropout", name="bneck_dropout", rate=dropout_rate))
    c = bn_f
    rev = list(reversed(FILTERS))
    for i, f in enumerate(rev):          # decoder, u_net.py:85-101
        s = len(rev) - i
        specs.append(dict(kind="convt", name=f"dec{s}_upsample", cin=c, cout=f))
        specs.append(dict(kind="concat", name=f"dec{s}_concat", skip=s))
        if dropout_rate > 0.0 and i < len(rev) - 1:
            specs.append(dict(kind="dropout", name=f"dec{s}_dropout", rate=dropout_rate))
        conv_block(f"dec{s}_block1", 2 * f, f)
        conv_block(f"dec{s}_block2", f, f)
        c = f
    specs.append(dict(kind="head", name="output_mask", cin=c, cout=num_classes,
                      activation="sigmoid" if num_classes == 1 else "softmax"))   # u_net.py:105-112
    return specs


def param_shapes(specs: List[dict]) -> "OrderedDict[str, Tuple[Tuple[int, ...], bool]]":
    """name -> (Keras shape, trainable), in Keras weight order per layer."""
    out: "OrderedDict[str, Tuple[Tuple[int, ...], bool]]" = OrderedDict()
    for sp in specs:
        n = sp["name"]
        if sp["kind"] == "sepconv":
            out[f"{n}/depthwise_kernel"] = ((3, 3, sp["cin"], 1), True)
            out[f"{n}/pointwise_kernel"] = ((1, 1, sp["cin"], sp["cout"]), True)
            if sp["use_bias"]:
                out[f"{n}/bias"] = ((sp["cout"],), True)
        elif sp["kind"] == "bn":
            out[f"{n}/gamma"] = ((sp["c"],), True)
            out[f"{n}/beta"] = ((sp["c"],), True)
            out[f"{n}/moving_mean"] = ((sp["c"],), False)
            out[f"{n}/moving_variance"] = ((sp["c"],), False)
        elif sp["kind"] == "convt":
            out[f"{n}/kernel"] = ((2, 2, sp["cout"], sp["cin"]), True)
            out[f"{n}/bias"] = ((sp["cout"],), True)
        elif sp["kind"] == "head":
            out[f"{n}/kernel"] = ((1, 1, sp["cin"], sp["cout"]), True)
            out[f"{n}/bias"] = ((sp["cout"],), True)
    return out


def count_params(specs: List[dict]) -> Tuple[int, int]:
    tr = sum(int(np.prod(s)) for s, t in param_shapes(specs).values() if t)
    nt = sum(int(np.prod(s)) for s, t in param_shapes(specs).values() if not t)
    return tr, nt


def init_params(specs: List[dict], seed: int = 2301, trained_like: bool = False) -> "OrderedDict[str, np.ndarray]":
    """Keras default initialisers restated: Glorot-uniform kernels (conv fans = receptive field x channels),
    zero biases, gamma 1, beta 0, moving_mean 0, moving_variance 1.  `trained_like` randomises the BN state so that
    folding and the beta/gamma paths are exercised."""
    rng = np.random.default_rng(seed)
    p: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for name, (shape, _) in param_shapes(specs).items():
        leaf = name.split("/")[-1]
        if leaf.endswith("kernel"):
            rf = int(np.prod(shape[:-2]))
            fan_in, fan_out = rf * shape[-2], rf * shape[-1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            p[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        elif leaf == "bias":
            p[name] = (rng.normal(0, 0.05, size=shape) if trained_like else np.zeros(shape)).astype(np.float32)
        elif leaf == "gamma":
            p[name] = (rng.uniform(0.5, 1.5, size=shape) if trained_like else np.ones(shape)).astype(np.float32)
        elif leaf == "beta":
            p[name] = (rng.normal(0, 0.1, size=shape) if trained_like else np.zeros(shape)).astype(np.float32)
        elif leaf == "moving_mean":
            p[name] = (rng.normal(0, 0.1, size=shape) if trained_like else np.zeros(shape)).astype(np.float32)
        elif leaf == "moving_variance":
            p[name] = (rng.uniform(0.5, 1.5, size=shape) if trained_like else np.ones(shape)).astype(np.float32)
        else:
            raise AssertionError(name)
    return p


# ----------------------------------------------------------------------------------------------- dropout mask
def dropout_words(group: np.ndarray, seed: int):
    """Restatement of dropout_words() in csrc/common.cuh: the two 32-bit words (a, b) a group of four consecutive elements
    draws from its index and the seed (uint32 arithmetic, wrapping)."""
    m32 = np.uint64(0xFFFFFFFF)
    g = group.astype(np.uint64) & m32
    s = (np.uint64(seed & 0xFFFFFFFF) * np.uint64(0x85EBCA6B) + np.uint64(0xC2B2AE35)) & m32
    x = ((g ^ s) * np.uint64(0x9E3779B1)) & m32
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x85EBCA77)) & m32
    x ^= x >> np.uint64(13)
    y = (x * np.uint64(0xC2B2AE3D)) & m32
    y ^= y >> np.uint64(16)
    return x.astype(np.uint32), y.astype(np.uint32)


def dropout_hash(idx: np.ndarray, seed: int) -> np.ndarray:
    """The 32-bit word that holds element idx's 16-bit field (dropout_hash() of csrc/common.cuh; unet_host_dropout_hash)."""
    idx = np.asarray(idx).astype(np.uint64)
    a, b = dropout_words(idx >> np.uint64(2), seed)
    return np.where((idx & np.uint64(2)) != 0, b, a).astype(np.uint32)


def dropout_fields(idx: np.ndarray, seed: int) -> np.ndarray:
    """16-bit field of every element: a.lo, a.hi, b.lo, b.hi for elements 4k .. 4k+3."""
    idx = np.asarray(idx).astype(np.uint64)
    h = dropout_hash(idx, seed)
    return np.where((idx & np.uint64(1)) == 1, h >> np.uint32(16), h & np.uint32(0xFFFF)).astype(np.uint32)


def dropout_multiplier(shape: Tuple[int, int, int, int], rate: float, seed: int, offset: int = 0) -> np.ndarray:
    """Mask * 1/(1-rate) over an NHWC tensor, indexed by the element's linear offset (Dropout, u_net.py:78,98).
    Restates dropout_mult() of csrc/common.cuh: element e keeps iff its 16-bit field is below floor(keep_prob * 65536).
    `offset`: linear offset of the first element (to evaluate a slab of a larger tensor)."""
    n = int(np.prod(shape))
    idx = np.uint64(offset) + np.arange(n, dtype=np.uint64)
    r = dropout_fields(idx, seed)
    keep = np.float32(1.0 - rate)
    thr = min(int(np.float32(keep) * np.float32(65536.0)), 65535)
    inv = np.float32(1.0) / keep
    return np.where(r < np.uint32(thr), inv, np.float32(0)).astype(np.float32).reshape(shape)


# ----------------------------------------------------------------------------------------------- primitive ops
def dwconv3x3(x, wd):
    """Depthwise half of SeparableConv2D: cross-correlation, zero 'same' padding.  wd: (3,3,C)."""
    n, h, w, c = x.shape
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    y = np.zeros_like(x)
    for a in range(3):
        for b in range(3):
            y += xp[:, a:a + h, b:b + w, :] * wd[a, b]
    return y


def dwconv3x3_bwd(x, wd, dy):
    n, h, w, c = x.shape
    dx = dwconv3x3(dy, wd[::-1, ::-1])
    xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
    dw = np.zeros_like(wd)
    for a in range(3):
        for b in range(3):
            dw[a, b] = np.sum(xp[:, a:a + h, b:b + w, :] * dy, axis=(0, 1, 2))
    return dx, dw


def maxpool2x2(x):
    n, h, w, c = x.shape
    xr = x[:, :h // 2 * 2, :w // 2 * 2].reshape(n, h // 2, 2, w // 2, 2, c)
    return xr.max(axis=(2, 4))


def maxpool2x2_bwd(x, dy):
    """Gradient goes to the first maximum of each window in scan order (TF CPU convention)."""
    n, h, w, c = x.shape
    xr = x.reshape(n, h // 2, 2, w // 2, 2, c).transpose(0, 1, 3, 5, 2, 4).reshape(n, h // 2, w // 2, c, 4)
    arg = np.argmax(xr, axis=-1)      # np.argmax returns the first maximum
    onehot = (arg[..., None] == np.arange(4)).astype(dy.dtype)
    d = onehot * dy[..., None]
    return d.reshape(n, h // 2, w // 2, c, 2, 2).transpose(0, 1, 4, 2, 5, 3).reshape(n, h, w, c)


def convt2x2(x, k, b):
    """Conv2DTranspose(f, 2, strides=2, 'same'): out[n,2i+a,2j+b,co] = sum_ci x[n,i,j,ci] k[a,b,co,ci] + bias."""
    n, h, w, ci = x.shape
    co = k.shape[2]
    y = np.einsum("nijc,abdc->niajbd", x, k, optimize=True).reshape(n, 2 * h, 2 * w, co)
    return y + b


def convt2x2_bwd(x, k, dy):
    n, h, w, ci = x.shape
    co = k.shape[2]
    g = dy.reshape(n, h, 2, w, 2, co)
    dx = np.einsum("niajbd,abdc->nijc", g, k, optimize=True)
    dk = np.einsum("niajbd,nijc->abdc", g, x, optimize=True)
    db = dy.sum(axis=(0, 1, 2))
    return dx, dk, db


def sigmoid(z):
    return 1.0 / (1.0 + np.exp(-z))


def softmax(z):
    z = z - z.max(axis=-1, keepdims=True)
    e = np.exp(z)
    return e / e.sum(axis=-1, keepdims=True)


# ----------------------------------------------------------------------------------------------- metrics / losses
def dice_coef(y_true, y_pred, smooth=EPSILON, dtype=np.float32):
    """utils/metrics.py:6-39, line by line."""
    y_true = np.asarray(y_true).astype(dtype)                      # :26
    y_pred = np.asarray(y_pred).astype(dtype)                      # :27
    inter = np.sum(y_true * y_pred, axis=(1, 2), dtype=dtype)      # :31
    st = np.sum(y_true, axis=(1, 2), dtype=dtype)                  # :32
    sp = np.sum(y_pred, axis=(1, 2), dtype=dtype)                  # :33
    num = dtype(2.0) * inter + dtype(smooth)                       # :35
    den = st + sp + dtype(smooth)                                  # :36
    return np.mean(num / den, dtype=dtype)                         # :37-38


def iou_coef(y_true, y_pred, smooth=EPSILON, dtype=np.float32):
    """utils/metrics.py:41-62."""
    y_true = np.asarray(y_true).astype(dtype)
    y_pred = np.asarray(y_pred).astype(dtype)
    inter = np.sum(y_true * y_pred, axis=(1, 2), dtype=dtype)
    st = np.sum(y_true, axis=(1, 2), dtype=dtype)
    sp = np.sum(y_pred, axis=(1, 2), dtype=dtype)
    union = st + sp - inter
    return np.mean((inter + dtype(smooth)) / (union + dtype(smooth)), dtype=dtype)


def dice_loss(y_true, y_pred, dtype=np.float32):
    """utils/loss.py:9-29."""
    return dtype(1.0) - dice_coef(y_true, y_pred, dtype=dtype)


def iou_loss(y_true, y_pred, smooth=EPSILON, dtype=np.float32):
    """utils/loss.py:31-45 (the reference forgets to import iou_coef, loss.py:4; the intended value is restated)."""
    return dtype(1.0) - iou_coef(y_true, y_pred, smooth=smooth, dtype=dtype)


jaccard_loss = iou_loss    # utils/loss.py:48


def sample_iou(y_true, y_pred, smooth=EPSILON):
    """scripts/benchmark.py:159-170: global sums over one squeezed sample."""
    t = np.asarray(y_true).squeeze().astype(np.float32)
    p = np.asarray(y_pred).squeeze().astype(np.float32)
    inter = np.sum(t * p, dtype=np.float32)
    union = np.sum(t, dtype=np.float32) + np.sum(p, dtype=np.float32) - inter
    return float((inter + np.float32(smooth)) / (union + np.float32(smooth)))


class MeanIoU:
    """tf.keras.metrics.MeanIoU(num_classes) (train.py:231, benchmark.py:237): labels and predictions are flattened
    and cast to integers by truncation; result is the mean of TP/(TP+FP+FN) over classes with a non-zero denominator."""

    def __init__(self, num_classes: int):
        self.num_classes = num_classes
        self.cm = np.zeros((num_classes, num_classes), dtype=np.int64)

    def reset_state(self):
        self.cm[:] = 0

    def update_state(self, y_true, y_pred):
        t = np.asarray(y_true).reshape(-1).astype(np.int64)       # truncation toward zero
        p = np.asarray(y_pred).reshape(-1).astype(np.int64)
        ok = (t >= 0) & (t < self.num_classes) & (p >= 0) & (p < self.num_classes)
        np.add.at(self.cm, (t[ok], p[ok]), 1)

    def result(self) -> float:
        cm = self.cm.astype(np.float64)
        tp = np.diag(cm)
        denom = cm.sum(axis=0) + cm.sum(axis=1) - tp
        valid = denom > 0
        if not valid.any():
            return 0.0
        iou = np.where(valid, tp / np.where(valid, denom, 1.0), 0.0)
        return float(iou.sum() / valid.sum())


def adamw_step(w, g, m, v, t, lr=2e-3, wd=1e-4, b1=0.9, b2=0.999, eps=1e-7):
    """Keras AdamW (train.py:226): decoupled decay on every variable, eps outside the bias correction. t starts at 1."""
    w = w - lr * wd * w
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    alpha = lr * np.sqrt(1 - b2 ** t) / (1 - b1 ** t)
    w = w - alpha * m / (np.sqrt(v) + eps)
    return w, m, v


# ----------------------------------------------------------------------------------------------- whole model
class UNetOracle:
    """Forward (+ analytic backward) of U_NET(input_size, num_classes, dropout_rate, use_batch_norm)."""

    def __init__(self, input_size, num_classes=1, dropout_rate=0.2, use_batch_norm=True, dtype=np.float64):
        self.specs = layer_specs(input_size, num_classes, dropout_rate, use_batch_norm)
        self.input_size = tuple(input_size)
        self.num_classes = num_classes
        self.dropout_rate = dropout_rate
        self.use_batch_norm = use_batch_norm
        self.dtype = dtype

    # -- helpers
    def _conv_block(self, x, prefix, P, training, cache, new_stats):
        dt = self.dtype
        wd = P[f"{prefix}_sepconv/depthwise_kernel"][..., 0].astype(dt)
        wp = P[f"{prefix}_sepconv/pointwise_kernel"][0, 0].astype(dt)
        d = dwconv3x3(x, wd)
        z = d.reshape(-1, wp.shape[0]) @ wp
        z = z.reshape(x.shape[:3] + (wp.shape[1],))
        c = dict(x=x, d=d, wd=wd, wp=wp)
        if self.use_batch_norm:
            g = P[f"{prefix}_bn/gamma"].astype(dt)
            b = P[f"{prefix}_bn/beta"].astype(dt)
            if training:
                mean = z.mean(axis=(0, 1, 2))
                var = z.var(axis=(0, 1, 2))                       # biased
                new_stats[f"{prefix}_bn/moving_mean"] = (P[f"{prefix}_bn/moving_mean"] * BN_MOMENTUM + mean * (1 - BN_MOMENTUM))
                new_stats[f"{prefix}_bn/moving_variance"] = (P[f"{prefix}_bn/moving_variance"] * BN_MOMENTUM + var * (1 - BN_MOMENTUM))
            else:
                mean = P[f"{prefix}_bn/moving_mean"].astype(dt)
                var = P[f"{prefix}_bn/moving_variance"].astype(dt)
            rstd = 1.0 / np.sqrt(var + dt(BN_EPS))
            xhat = (z - mean) * rstd
            pre = xhat * g + b
            c.update(xhat=xhat, rstd=rstd, gamma=g)
        else:
            pre = z + P[f"{prefix}_sepconv/bias"].astype(dt)
        y = np.maximum(pre, 0)
        c["y"] = y
        cache[prefix] = c
        return y

    def forward(self, P: Dict[str, np.ndarray], x: np.ndarray, training: bool = False,
                drop_seeds: Optional[Dict[str, int]] = None, keep_cache: bool = False):
        """Returns probabilities (N,H,W,C); with keep_cache also the tape and the new BN moving statistics."""
        dt = self.dtype
        x = np.asarray(x).astype(dt)
        cache: Dict[str, dict] = {}
        new_stats: Dict[str, np.ndarray] = {}
        drop_seeds = drop_seeds or {}

        def dropout(name, t):
            if not training or self.dropout_rate <= 0.0:
                return t
            mult = dropout_multiplier(t.shape, self.dropout_rate, drop_seeds[name]).astype(dt)
            cache[name] = dict(mult=mult)
            return t * mult

        skips = []
        for s in range(1, 5):
            x = self._conv_block(x, f"enc{s}_block1", P, training, cache, new_stats)
            x = self._conv_block(x, f"enc{s}_block2", P, training, cache, new_stats)
            skips.append(x)
            cache[f"enc{s}_pool"] = dict(x=x)
            x = maxpool2x2(x)
        x = self._conv_block(x, "bneck_block1", P, training, cache, new_stats)
        x = self._conv_block(x, "bneck_block2", P, training, cache, new_stats)
        x = dropout("bneck_dropout", x)
        for i, s in enumerate([4, 3, 2, 1]):
            k = P[f"dec{s}_upsample/kernel"].astype(dt)
            b = P[f"dec{s}_upsample/bias"].astype(dt)
            cache[f"dec{s}_upsample"] = dict(x=x, k=k)
            up = convt2x2(x, k, b)
            x = np.concatenate([up, skips[s - 1]], axis=-1)        # [upsampled, skip], u_net.py:96
            if i < 3:
                x = dropout(f"dec{s}_dropout", x)
            x = self._conv_block(x, f"dec{s}_block1", P, training, cache, new_stats)
            x = self._conv_block(x, f"dec{s}_block2", P, training, cache, new_stats)
        wk = P["output_mask/kernel"][0, 0].astype(dt)
        bk = P["output_mask/bias"].astype(dt)
        logits = x.reshape(-1, wk.shape[0]) @ wk + bk
        logits = logits.reshape(x.shape[:3] + (self.num_classes,))
        probs = sigmoid(logits) if self.num_classes == 1 else softmax(logits)
        cache["output_mask"] = dict(x=x, wk=wk, probs=probs)
        if keep_cache:
            return probs, cache, new_stats
        return probs

    # -- backward
    def _conv_block_bwd(self, prefix, dy, cache, grads, need_dx=True):
        c = cache[prefix]
        g = dy * (c["y"] > 0)
        if self.use_batch_norm:
            xhat, rstd, gamma = c["xhat"], c["rstd"], c["gamma"]
            grads[f"{prefix}_bn/gamma"] = np.sum(g * xhat, axis=(0, 1, 2))
            grads[f"{prefix}_bn/beta"] = np.sum(g, axis=(0, 1, 2))
            m = g.shape[0] * g.shape[1] * g.shape[2]
            dz = gamma * rstd * (g - grads[f"{prefix}_bn/beta"] / m - xhat * grads[f"{prefix}_bn/gamma"] / m)
        else:
            grads[f"{prefix}_sepconv/bias"] = np.sum(g, axis=(0, 1, 2))
            dz = g
        d, wp, wd, x = c["d"], c["wp"], c["wd"], c["x"]
        dz2 = dz.reshape(-1, wp.shape[1])
        grads[f"{prefix}_sepconv/pointwise_kernel"] = (d.reshape(-1, wp.shape[0]).T @ dz2)[None, None]
        dd = (dz2 @ wp.T).reshape(d.shape)
        dx, dwd = dwconv3x3_bwd(x, wd, dd)
        grads[f"{prefix}_sepconv/depthwise_kernel"] = dwd[..., None]
        return dx if need_dx else None

    def loss_and_grads(self, P, x, y_true, loss: str = "dice", drop_seeds=None):
        """Training-mode forward, loss (utils/loss.py), and d loss / d every trainable parameter."""
        dt = self.dtype
        probs, cache, new_stats = self.forward(P, x, training=True, drop_seeds=drop_seeds, keep_cache=True)
        t = np.asarray(y_true).astype(dt)
        n, h, w, c = probs.shape
        inter = np.sum(t * probs, axis=(1, 2)); st = np.sum(t, axis=(1, 2)); sp = np.sum(probs, axis=(1, 2))
        if loss == "dice":
            den = st + sp + EPSILON
            num = 2 * inter + EPSILON
            value = 1.0 - np.mean(num / den)
            dp = -(1.0 / (n * c)) * (2 * t * den[:, None, None, :] - num[:, None, None, :]) / (den ** 2)[:, None, None, :]
        elif loss == "iou":
            u = st + sp - inter + EPSILON
            i_s = inter + EPSILON
            value = 1.0 - np.mean(i_s / u)
            dp = -(1.0 / (n * c)) * (t * u[:, None, None, :] - i_s[:, None, None, :] * (1 - t)) / (u ** 2)[:, None, None, :]
        else:
            raise ValueError(loss)
        grads: Dict[str, np.ndarray] = {}
        if self.num_classes == 1:
            dlogit = dp * probs * (1 - probs)
        else:
            dlogit = probs * (dp - np.sum(dp * probs, axis=-1, keepdims=True))
        hc = cache["output_mask"]
        xk = hc["x"]
        dl2 = dlogit.reshape(-1, c)
        grads["output_mask/kernel"] = (xk.reshape(-1, xk.shape[-1]).T @ dl2)[None, None]
        grads["output_mask/bias"] = dl2.sum(axis=0)
        dx = (dl2 @ hc["wk"].T).reshape(xk.shape)
        dskips = {}
        for i, s in enumerate([1, 2, 3, 4]):
            dx = self._conv_block_bwd(f"dec{s}_block2", dx, cache, grads)
            dx = self._conv_block_bwd(f"dec{s}_block1", dx, cache, grads)
            if f"dec{s}_dropout" in cache:
                dx = dx * cache[f"dec{s}_dropout"]["mult"]
            f = dx.shape[-1] // 2
            dup, dskips[s] = dx[..., :f], dx[..., f:]
            uc = cache[f"dec{s}_upsample"]
            dx, dk, db = convt2x2_bwd(uc["x"], uc["k"], dup)
            grads[f"dec{s}_upsample/kernel"] = dk
            grads[f"dec{s}_upsample/bias"] = db
        if "bneck_dropout" in cache:
            dx = dx * cache["bneck_dropout"]["mult"]
        dx = self._conv_block_bwd("bneck_block2", dx, cache, grads)
        dx = self._conv_block_bwd("bneck_block1", dx, cache, grads)
        for s in [4, 3, 2, 1]:
            dx = maxpool2x2_bwd(cache[f"enc{s}_pool"]["x"], dx) + dskips[s]
            dx = self._conv_block_bwd(f"enc{s}_block2", dx, cache, grads)
            dx = self._conv_block_bwd(f"enc{s}_block1", dx, cache, grads, need_dx=(s != 1))
        return float(value), probs, grads, new_stats


def synthetic_batch(n: int, h: int, w: int, cin: int = 3, num_classes: int = 1, seed: int = 2301):
    """Images uniform[0,1); binary masks = a filled random quadrilateral-ish half-plane intersection per image
    (mirrors the quad masks of benchmark.py:135-148); multi-class = labels piecewise-constant on a coarse grid, one-hot."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, h, w, cin), dtype=np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    if num_classes == 1:
        y = np.zeros((n, h, w, 1), np.float32)
        for i in range(n):
            cx, cy = rng.uniform(0.3, 0.7) * w, rng.uniform(0.3, 0.7) * h
            rx, ry = rng.uniform(0.15, 0.35) * w, rng.uniform(0.15, 0.35) * h
            th = rng.uniform(0, np.pi)
            u = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
            v = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
            y[i, ..., 0] = ((np.abs(u) < rx) & (np.abs(v) < ry)).astype(np.float32)
        return x, y
    g = max(1, min(h, w) // 8)
    lab = rng.integers(0, num_classes, size=(n, (h + g - 1) // g, (w + g - 1) // g))
    lab = np.repeat(np.repeat(lab, g, axis=1), g, axis=2)[:, :h, :w]
    y = (lab[..., None] == np.arange(num_classes)).astype(np.float32)
    return x, y
