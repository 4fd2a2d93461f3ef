"""CPU pre/post-processing shared by the CLIs, restating what the reference scripts do with OpenCV around
`model.predict` (scripts/inference.py:98-110,127-201; scripts/benchmark.py:95-170).  Not on the GPU hot path."""
from __future__ import annotations

import json
import os
from typing import Optional, Tuple

import numpy as np

SMOOTH = 1e-7


def read_image_for_model(path: str, height: int, width: int):
    """cv2.imread (BGR) -> float32 / 255 -> bilinear resize -> (1, H, W, 3).  BGR order is the reference's behaviour
    (inference.py:100-109), kept as is.  Returns (tensor, original BGR image) or (None, None)."""
    import cv2
    bgr = cv2.imread(path, cv2.IMREAD_COLOR)
    if bgr is None:
        return None, None
    x = cv2.resize(bgr.astype(np.float32) / 255.0, (width, height), interpolation=cv2.INTER_LINEAR)
    return x[None], bgr


def probability_to_mask(prob: np.ndarray, out_height: int, out_width: int, threshold: float) -> np.ndarray:
    """(h, w, 1) probabilities -> bilinear resize to the original size -> uint8 {0, 255} (inference.py:147-160)."""
    import cv2
    p = cv2.resize(prob, (out_width, out_height), interpolation=cv2.INTER_LINEAR)
    if p.ndim == 3:
        p = p[..., 0]
    return (p > threshold).astype(np.uint8) * 255


def largest_region_crop(mask: np.ndarray, image_bgr: np.ndarray, min_area: float):
    """Bounding box of the largest external contour (by cv2.contourArea) if its area exceeds `min_area`
    (inference.py:173-187).  Returns (crop or None, area or None, rect or None)."""
    import cv2
    contours, _ = cv2.findContours(mask, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    if not contours:
        return None, None, None
    best = max(contours, key=cv2.contourArea)
    area = cv2.contourArea(best)
    if area <= min_area:
        return None, area, None
    x, y, w, h = cv2.boundingRect(best)
    return image_bgr[y:y + h, x:x + w], area, (x, y, w, h)


def quad_mask(json_path: str, height: int, width: int) -> Optional[np.ndarray]:
    """Ground-truth mask of scripts/benchmark.py:112-157: fill the JSON `quad` polygon on a canvas of the companion
    image's size (2048x2048 if the image is missing), nearest-resize to (height, width), > 128 -> (1, H, W, 1) uint8."""
    import cv2
    with open(json_path) as f:
        quad = json.load(f).get("quad", [])
    oh = ow = -1
    for ext in (".tif", ".png", ".jpg"):
        cand = json_path.replace("/ground_truth/", "/images/").replace(".json", ext)
        if os.path.exists(cand):
            img = cv2.imread(cand, cv2.IMREAD_UNCHANGED)
            if img is not None:
                oh, ow = img.shape[:2]
                break
    if oh <= 0 or ow <= 0:
        print(f"Warning: Could not determine original dimensions for mask from {json_path}. "
              "Using default large canvas (2048x2048).")
        oh = ow = 2048
    canvas = np.zeros((oh, ow), np.uint8)
    if quad:
        pts = np.array(quad, dtype=np.int32)
        if pts.ndim == 2:
            pts = pts.reshape(-1, 1, 2)
        cv2.drawContours(canvas, [pts], contourIdx=-1, color=255, thickness=cv2.FILLED)
    small = cv2.resize(canvas, (width, height), interpolation=cv2.INTER_NEAREST)
    return (small > 128).astype(np.uint8)[None, ..., None]


def sample_iou(y_true: np.ndarray, y_pred: np.ndarray, smooth: float = SMOOTH) -> float:
    """Global (I+s)/(T+P-I+s) of one sample (scripts/benchmark.py:159-170)."""
    t = np.asarray(y_true, np.float32).squeeze()
    p = np.asarray(y_pred, np.float32).squeeze()
    inter = np.float32((t * p).sum(dtype=np.float32))
    union = np.float32(t.sum(dtype=np.float32)) + np.float32(p.sum(dtype=np.float32)) - inter
    return float((inter + np.float32(smooth)) / (union + np.float32(smooth)))


def sample_iou_from_counts(cm4, smooth: float = SMOOTH) -> float:
    """calculate_sample_iou (scripts/benchmark.py:159-170) from a sample's 2x2 confusion counts [t*2+p]: for {0,1} arrays
    I = #(t=1,p=1), T = #(t=1), P = #(p=1); same float32 arithmetic as sample_iou, so the value is identical."""
    inter = np.float32(cm4[3])
    union = np.float32(cm4[2] + cm4[3]) + np.float32(cm4[1] + cm4[3]) - inter
    return float((inter + np.float32(smooth)) / (union + np.float32(smooth)))


# ------------------------------------------------------------------------------------------------ GPU variants (opt-in)
def gpu_preprocess(images_bgr, height: int, width: int):
    """uint8 BGR images (any sizes) -> one CUDA fp32 batch [B, height, width, 3]: the same /255 + INTER_LINEAR resize as
    read_image_for_model, done by unet_preprocess_u8 (one upload of the raw bytes per image instead of 4x as many floats)."""
    import torch
    from . import ops
    out = torch.empty((len(images_bgr), height, width, 3), dtype=torch.float32, device="cuda")
    for i, img in enumerate(images_bgr):
        if img.dtype != np.uint8 or img.ndim != 3 or img.shape[2] != 3:
            raise ValueError("gpu_preprocess expects uint8 HxWx3 images")
        ops.preprocess_u8(torch.from_numpy(np.ascontiguousarray(img)).cuda(), out[i])
    return out


def gpu_probability_to_mask(prob_dev, out_height: int, out_width: int, threshold: float) -> np.ndarray:
    """CUDA probabilities [h, w, 1] (a slice of the model output) -> uint8 {0,255} mask at the original size on the host."""
    import torch
    from . import ops
    mask = torch.empty((out_height, out_width), dtype=torch.uint8, device="cuda")
    ops.postprocess_mask(prob_dev[..., 0], mask, threshold)
    return mask.cpu().numpy()


def gpu_batch_sample_counts(truth, prob_dev, threshold: float) -> np.ndarray:
    """truth: host uint8/float {0,1} [B,h,w,1]; prob_dev: CUDA fp32 [B,h,w,1] (model.predict_on_device).  Returns the
    per-sample 2x2 confusion counts [B,4] (int64, host) of (prob > threshold) — one kernel launch, one 32*B-byte read-back
    instead of the probabilities' trip to the host."""
    import torch
    from . import ops
    t = torch.from_numpy(np.ascontiguousarray(truth, dtype=np.float32)).cuda()
    counts = torch.zeros((t.shape[0], 4), dtype=torch.int64, device="cuda")
    ops.sample_confusion_thr(t, prob_dev, threshold, counts)
    return counts.cpu().numpy()
