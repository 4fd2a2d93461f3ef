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
    """Infinite iterator over (frames, masks) batches from two directories with identically sorted file lists.

    `workers` > 0: images are decoded / resized by a thread pool (OpenCV releases the GIL) and whole batches are prepared
    `prefetch` deep by a background thread, so the host keeps pace with a GPU step that consumes ~1000 images/s; the
    shuffle order and the flip decisions are drawn by one thread in batch order, so the stream is identical to the
    synchronous path (`workers=0`) for the same seed."""

    def __init__(self, frames_dir: str, masks_dir: str, target_size=(256, 256), batch_size=2, shuffle=True,
                 horizontal_flip=False, rescale=1.0 / 255.0, seed: Optional[int] = None,
                 workers: Optional[int] = None, prefetch: int = 2):
        self.frames, self.masks = list_images(frames_dir), list_images(masks_dir)
        if len(self.frames) != len(self.masks):
            raise ValueError(f"{len(self.frames)} frames vs {len(self.masks)} masks")
        self.n = len(self.frames)
        self.samples = self.n
        self.target_size, self.batch_size = tuple(target_size), int(batch_size)
        self.shuffle, self.flip, self.rescale = shuffle, horizontal_flip, float(rescale)
        self.rng = np.random.default_rng(seed)
        self.workers = min(16, os.cpu_count() or 1) if workers is None else max(0, int(workers))
        self.prefetch = max(1, int(prefetch))
        print(f"Found {self.n} images belonging to 1 classes.")

    # ------------------------------------------------------------------ one sample / one batch
    def _fill(self, x: np.ndarray, y: np.ndarray, k: int, i: int, flip: bool) -> None:
        xi = _load(self.frames[i], self.target_size, gray=False)
        yi = _load(self.masks[i], self.target_size, gray=True)
        if flip:
            xi, yi = xi[:, ::-1], yi[:, ::-1]
        scale = np.float32(self.rescale)        # float32 image * float32(1/255), as Keras' `x *= rescale` on img_to_array output
        np.multiply(xi, scale, out=x[k], dtype=np.float32, casting="unsafe")
        np.multiply(yi, scale, out=y[k], dtype=np.float32, casting="unsafe")

    def _plan(self) -> Iterator[Tuple[np.ndarray, List[bool]]]:
        """Batch plans (sample indices, flip flags) in stream order: the only consumer of the RNG."""
        while True:
            order = self.rng.permutation(self.n) if self.shuffle else np.arange(self.n)
            for lo in range(0, self.n, self.batch_size):
                idx = order[lo:lo + self.batch_size]
                flips = [bool(self.flip and self.rng.random() < 0.5) for _ in idx]
                yield idx, flips

    def _batch(self, idx, flips, pool=None) -> Tuple[np.ndarray, np.ndarray]:
        h, w = self.target_size
        x = np.empty((len(idx), h, w, 3), np.float32)
        y = np.empty((len(idx), h, w, 1), np.float32)
        if pool is None:
            for k, (i, f) in enumerate(zip(idx, flips)):
                self._fill(x, y, k, int(i), f)
        else:
            for fut in [pool.submit(self._fill, x, y, k, int(i), f) for k, (i, f) in enumerate(zip(idx, flips))]:
                fut.result()                    # re-raises a decode error in the consumer's stack
        return x, y

    def __iter__(self) -> Iterator[Tuple[np.ndarray, np.ndarray]]:
        if self.workers == 0:
            for idx, flips in self._plan():
                yield self._batch(idx, flips)
            return
        import queue
        import threading
        from concurrent.futures import ThreadPoolExecutor
        ready: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()
        pool = ThreadPoolExecutor(self.workers, thread_name_prefix="unet_decode")

        def producer():
            try:
                for idx, flips in self._plan():
                    item = self._batch(idx, flips, pool)
                    while not stop.is_set():
                        try:
                            ready.put(item, timeout=0.1)
                            break
                        except queue.Full:
                            continue
                    if stop.is_set():
                        return
            except BaseException as exc:        # handed to the consumer, which re-raises it
                while not stop.is_set():
                    try:
                        ready.put(exc, timeout=0.1)
                        return
                    except queue.Full:
                        continue

        th = threading.Thread(target=producer, name="unet_batches", daemon=True)
        th.start()
        try:
            while True:
                item = ready.get()
                if isinstance(item, BaseException):
                    raise item
                yield item
        finally:                                # generator closed or collected: stop the producer, free the pool
            stop.set()
            th.join(timeout=5.0)
            pool.shutdown(wait=False)


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
