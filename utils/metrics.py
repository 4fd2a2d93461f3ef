"""Drop-in for the reference `utils/metrics.py` (dice_coef :6-39, iou_coef :41-62).

Same definitions — per (sample, channel) sums over the spatial axes, (2I+s)/(T+P+s) resp. (I+s)/(T+P-I+s), mean
over (batch, channel) — computed by the `unet_seg_sums` / `unet_seg_loss_finalize` kernels of libunet_b200.so.
Inputs: array-likes or torch tensors of shape (batch, H, W, C); returns a host scalar that also has `.numpy()`.
"""
from __future__ import annotations

import numpy as np

SMOOTH = 1e-7   # K.epsilon()


def _sums(y_true, y_pred):
    import torch
    from unet_b200 import ops
    from unet_b200.keras_api import _to_device_f32
    t, p = _to_device_f32(y_true), _to_device_f32(y_pred)
    if t.dim() != 4 or tuple(t.shape) != tuple(p.shape):
        raise ValueError(f"expected y_true and y_pred of equal shape (batch, H, W, C); got {tuple(t.shape)}, {tuple(p.shape)}")
    nb, c = t.shape[0], t.shape[3]
    sums = torch.zeros((nb, c, 3), device="cuda", dtype=torch.float64)
    ops.seg_sums(t, p, sums)
    return sums, nb * c


def _coef(y_true, y_pred, smooth, which: int):
    import torch
    from unet_b200 import ops
    from unet_b200.keras_api import Scalar
    sums, npairs = _sums(y_true, y_pred)
    out3 = torch.empty(3, device="cuda")
    ops.seg_loss_finalize(sums, npairs, float(smooth), 0, 1.0, out3, None)
    return Scalar(float(out3[which]))


def dice_coef(y_true, y_pred, smooth: float = SMOOTH):
    """Dice coefficient averaged over batch and channels (reference utils/metrics.py:6-39)."""
    return _coef(y_true, y_pred, smooth, 1)


def iou_coef(y_true, y_pred, smooth: float = SMOOTH):
    """IoU (Jaccard) coefficient averaged over batch and channels (reference utils/metrics.py:41-62)."""
    return _coef(y_true, y_pred, smooth, 2)
