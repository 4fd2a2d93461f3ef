"""CPU oracle (TEST INFRASTRUCTURE ONLY) — second, independent restatement of the reference forward pass in
torch-CPU ops with autograd.  Used (a) to cross-check the NumPy oracle's analytic gradients, and (b) as the
multi-threaded CPU stand-in for the reference's TensorFlow CPU path in bench.py (`cpu_baseline`, `--impl reference`):
TensorFlow is not installable in this image, so the timed CPU arm is this port ("kind": "port").

PARITY UNPINNED — see oracle/unet_ref.py.  Follows model/u_net.py:5-116 for topology and SURVEY.md Appendix A for
the Keras layer semantics (torch mappings: depthwise = conv2d(groups=C) with kernel permuted (2,3,0,1);
Conv2DTranspose kernel (2,2,Cout,Cin) -> conv_transpose2d weight permuted (3,2,0,1); BN eps 1e-3, biased variance).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from .unet_ref import BN_EPS, EPSILON, dropout_multiplier


def to_torch(P: Dict[str, np.ndarray], dtype=torch.float64, requires_grad=False) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in P.items():
        t = torch.tensor(np.asarray(v), dtype=dtype)
        trainable = not (k.endswith("moving_mean") or k.endswith("moving_variance"))
        if requires_grad and trainable:
            t.requires_grad_(True)
        out[k] = t
    return out


def _conv_block(x, prefix, P, training, use_bn):
    wd = P[f"{prefix}_sepconv/depthwise_kernel"]          # (3,3,Cin,1)
    wp = P[f"{prefix}_sepconv/pointwise_kernel"]          # (1,1,Cin,Cout)
    cin = wd.shape[2]
    x = F.conv2d(x, wd.permute(2, 3, 0, 1), padding=1, groups=cin)
    bias = None if use_bn else P[f"{prefix}_sepconv/bias"]
    x = F.conv2d(x, wp.permute(3, 2, 0, 1), bias=bias)
    if use_bn:
        g, b = P[f"{prefix}_bn/gamma"], P[f"{prefix}_bn/beta"]
        if training:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
        else:
            mean, var = P[f"{prefix}_bn/moving_mean"], P[f"{prefix}_bn/moving_variance"]
        x = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + BN_EPS)
        x = x * g[None, :, None, None] + b[None, :, None, None]
    return F.relu(x)


def forward(P: Dict[str, torch.Tensor], x_nhwc: torch.Tensor, num_classes: int = 1, dropout_rate: float = 0.2,
            use_batch_norm: bool = True, training: bool = False, drop_seeds: Optional[Dict[str, int]] = None):
    """x: (N,H,W,C) -> probabilities (N,H,W,num_classes)."""
    x = x_nhwc.permute(0, 3, 1, 2)

    def dropout(name, t):
        if not training or dropout_rate <= 0.0:
            return t
        if drop_seeds is None:          # timing runs: torch's own Bernoulli mask (same cost class as TF's)
            return F.dropout(t, dropout_rate, training=True)
        n, c, h, w = t.shape
        mult = dropout_multiplier((n, h, w, c), dropout_rate, drop_seeds[name])
        return t * torch.tensor(mult, dtype=t.dtype).permute(0, 3, 1, 2)

    skips = []
    for s in range(1, 5):
        x = _conv_block(x, f"enc{s}_block1", P, training, use_batch_norm)
        x = _conv_block(x, f"enc{s}_block2", P, training, use_batch_norm)
        skips.append(x)
        x = F.max_pool2d(x, 2)
    x = _conv_block(x, "bneck_block1", P, training, use_batch_norm)
    x = _conv_block(x, "bneck_block2", P, training, use_batch_norm)
    x = dropout("bneck_dropout", x)
    for i, s in enumerate([4, 3, 2, 1]):
        k = P[f"dec{s}_upsample/kernel"]                  # (2,2,Cout,Cin)
        x = F.conv_transpose2d(x, k.permute(3, 2, 0, 1), bias=P[f"dec{s}_upsample/bias"], stride=2)
        x = torch.cat([x, skips[s - 1]], dim=1)
        if i < 3:
            x = dropout(f"dec{s}_dropout", x)
        x = _conv_block(x, f"dec{s}_block1", P, training, use_batch_norm)
        x = _conv_block(x, f"dec{s}_block2", P, training, use_batch_norm)
    wk = P["output_mask/kernel"]
    x = F.conv2d(x, wk.permute(3, 2, 0, 1), bias=P["output_mask/bias"])
    x = torch.sigmoid(x) if num_classes == 1 else torch.softmax(x, dim=1)
    return x.permute(0, 2, 3, 1)


def dice_coef(y_true, y_pred, smooth=EPSILON):
    inter = (y_true * y_pred).sum(dim=(1, 2))
    return ((2.0 * inter + smooth) / (y_true.sum(dim=(1, 2)) + y_pred.sum(dim=(1, 2)) + smooth)).mean()


def iou_coef(y_true, y_pred, smooth=EPSILON):
    inter = (y_true * y_pred).sum(dim=(1, 2))
    union = y_true.sum(dim=(1, 2)) + y_pred.sum(dim=(1, 2)) - inter
    return ((inter + smooth) / (union + smooth)).mean()


def loss_and_grads(P_np, x_np, y_np, num_classes=1, dropout_rate=0.2, use_batch_norm=True, loss="dice",
                   drop_seeds=None, dtype=torch.float64):
    P = to_torch(P_np, dtype=dtype, requires_grad=True)
    x = torch.tensor(x_np, dtype=dtype)
    y = torch.tensor(y_np, dtype=dtype)
    probs = forward(P, x, num_classes, dropout_rate, use_batch_norm, training=True, drop_seeds=drop_seeds)
    value = 1.0 - (dice_coef(y, probs) if loss == "dice" else iou_coef(y, probs))
    value.backward()
    grads = {k: v.grad.numpy() for k, v in P.items() if v.requires_grad and v.grad is not None}
    return float(value), probs.detach().numpy(), grads
