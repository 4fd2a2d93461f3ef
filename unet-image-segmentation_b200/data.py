"""Input pipeline equivalent to the reference's `ImageDataGenerator(...).flow_from_directory(...)` pairs
(scripts/train.py:169-220): images RGB bilinear / masks grayscale nearest, rescale 1/255, synchronized
horizontal flip, seeded shuffle, infinite generator of (x, y) float32 NHWC batches; the last partial batch of an epoch is
yielded smaller, as Keras does.  The RNG stream is our own (Keras' is unpinned): distributionally equivalent."""
from __future__ import annotations

import os
from typing import Iterator, List, Optional, Tuple

import numpy as np

IMG_EXT = (".png", ".jpg", ".jpeg", ".bmp", ".ppm", ".tif", ".tiff")


def list_images(directory: str) -> List[str]:
    """Files under `directory` (flow_from_directory is pointed at the parent with classes=['image'])."""
    out = []
    for root, _, files in sorted(os.walk(directory)):
        for f in sorted(files):
            if f.lower().endswith(IMG_EXT):
                out.append(os.path.join(root, f))
    return out


def _load(path: str, size: Tuple[int, int], gray: bool) -> np.ndarray:
    import cv2
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE if gray else cv2.IMREAD_COLOR)
    if img is None:
        raise OSError(f"cannot read image {path}")
    if not gray:
        img = cv2.cvtColor(img, cv2.COLOR_BGR2RGB)
    img = cv2.resize(img, (size[1], size[0]), interpolation=cv2.INTER_NEAREST if gray else cv2.INTER_LINEAR)
    return img[..., None] if gray else img


class PairedDirectoryIterator:
    """Infinite iterator over (frames, masks) batches from two directories with identically sorted file lists."""

    def __init__(self, frames_dir: str, masks_dir: str, target_size=(256, 256), batch_size=2, shuffle=True,
                 horizontal_flip=False, rescale=1.0 / 255.0, seed: Optional[int] = None):
        self.frames, self.masks = list_images(frames_dir), list_images(masks_dir)
        if len(self.frames) != len(self.masks):
            raise ValueError(f"{len(self.frames)} frames vs {len(self.masks)} masks")
        self.n = len(self.frames)
        self.samples = self.n
        self.target_size, self.batch_size = tuple(target_size), int(batch_size)
        self.shuffle, self.flip, self.rescale = shuffle, horizontal_flip, float(rescale)
        self.rng = np.random.default_rng(seed)
        print(f"Found {self.n} images belonging to 1 classes.")

    def __iter__(self) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
        while True:
            order = self.rng.permutation(self.n) if self.shuffle else np.arange(self.n)
            for lo in range(0, self.n, self.batch_size):
                idx = order[lo:lo + self.batch_size]
                h, w = self.target_size
                x = np.empty((len(idx), h, w, 3), np.float32)
                y = np.empty((len(idx), h, w, 1), np.float32)
                for k, i in enumerate(idx):
                    xi = _load(self.frames[i], self.target_size, gray=False).astype(np.float32) * self.rescale
                    yi = _load(self.masks[i], self.target_size, gray=True).astype(np.float32) * self.rescale
                    if self.flip and self.rng.random() < 0.5:
                        xi, yi = xi[:, ::-1], yi[:, ::-1]
                    x[k], y[k] = xi, yi
                yield x, y


def synthetic_batches(batch: int, height: int, width: int, classes: int = 1, seed: int = 2301,
                      pool: int = 8) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
    """Synthetic (x, y) stream of the named shape: uniform images, filled-rectangle masks (multi-class: one-hot of
    a coarse label grid).  A small pool of pre-built batches is cycled."""
    rng = np.random.default_rng(seed)
    items = []
    yy, xx = np.mgrid[0:height, 0:width]
    for _ in range(pool):
        x = rng.random((batch, height, width, 3), dtype=np.float32)
        if classes == 1:
            y = np.zeros((batch, height, width, 1), np.float32)
            for i in range(batch):
                cx, cy = rng.uniform(0.3, 0.7) * width, rng.uniform(0.3, 0.7) * height
                rx, ry = rng.uniform(0.15, 0.35) * width, rng.uniform(0.15, 0.35) * height
                y[i, ..., 0] = ((np.abs(xx - cx) < rx) & (np.abs(yy - cy) < ry)).astype(np.float32)
        else:
            g = max(1, min(height, width) // 8)
            lab = rng.integers(0, classes, size=(batch, (height + g - 1) // g, (width + g - 1) // g))
            lab = np.repeat(np.repeat(lab, g, 1), g, 2)[:, :height, :width]
            y = (lab[..., None] == np.arange(classes)).astype(np.float32)
        items.append((x, y))
    while True:
        for it in items:
            yield it
