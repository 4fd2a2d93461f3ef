#!/usr/bin/env python3
"""benchmark.py — accuracy evaluation on the B200 engine, drop-in for the reference's scripts/benchmark.py.

CLI surface kept from the reference (scripts/benchmark.py:59-93): positional `input_dir`, `--model` (./models/model.h5),
`--iou_threshold` (0.9), `--pred_threshold` (0.5), `--low_score_log`; exit code 1 on any invalid argument.
Flow (reference :95-290): every images/**/*.tif is paired with ground_truth/**.json, the JSON quadrilateral becomes the
true mask, the thresholded prediction is scored per sample (global IoU) and accumulated into MeanIoU(num_classes=2);
samples below `--iou_threshold` are listed and optionally written as CSV `FileID,MeanIoU_Score`.
Extension: `--batch N` evaluates N images per `model.predict` call (the reference runs batch 1; scores are identical).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from glob import glob
from typing import List, Tuple

import numpy as np

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), os.pardir))
if ROOT not in sys.path:
    sys.path.append(ROOT)

# (flags, kwargs) — names and defaults are the reference's
_ARGS = [
    (("input_dir",), dict(type=str, help="dataset root holding images/ and ground_truth/")),
    (("--model",), dict(type=str, default="./models/model.h5", help="model file (.h5 / .keras / .npz)")),
    (("--iou_threshold",), dict(type=float, default=0.9, help="samples scoring below this IoU are reported")),
    (("--pred_threshold",), dict(type=float, default=0.5, help="probability above which a pixel counts as foreground")),
    (("--low_score_log",), dict(type=str, default=None, help="CSV file for the samples below --iou_threshold")),
    (("--batch",), dict(type=int, default=1, help="images per predict call (extension; the reference uses 1)")),
    (("--host-metrics",), dict(action="store_true", help="threshold and score on the host with NumPy, as the reference does "
                                                         "(default: integer confusion counts on the GPU; identical numbers)")),
]


def parse_args(argv=None) -> argparse.Namespace:
    parser = argparse.ArgumentParser(description="Score a U-Net segmentation model against JSON quadrilateral ground truth.")
    for flags, kw in _ARGS:
        parser.add_argument(*flags, **kw)
    return parser.parse_args(argv)


def _die(message: str):
    print(f"Error: {message}")
    sys.exit(1)


def _validate(args) -> Tuple[str, str]:
    """Argument checks of the reference (scripts/benchmark.py:179-193): each failure prints and exits with code 1."""
    images_root, truth_root = (os.path.join(args.input_dir, sub) for sub in ("images", "ground_truth"))
    checks = [
        (os.path.isdir(args.input_dir), f"no such input directory: {args.input_dir}"),
        (os.path.isdir(images_root) and os.path.isdir(truth_root), f"expected sub-directories '{images_root}' and '{truth_root}'"),
        (os.path.isfile(args.model), f"no such model file: {args.model}"),
        (0.0 <= args.pred_threshold <= 1.0, f"--pred_threshold outside [0, 1]: {args.pred_threshold}"),
        (0.0 <= args.iou_threshold <= 1.0, f"--iou_threshold outside [0, 1]: {args.iou_threshold}"),
    ]
    for ok, message in checks:
        if not ok:
            _die(message)
    return images_root, truth_root


def _collect_pairs(images_root: str, truth_root: str) -> List[Tuple[str, str, str]]:
    """(image path, JSON path, id) for every .tif that has a ground-truth file at the same relative path."""
    tifs = sorted(glob(os.path.join(images_root, "**", "*.tif"), recursive=True))
    print(f"{len(tifs)} .tif images under {images_root}")
    pairs, orphans = [], 0
    for tif in tifs:
        sample_id = os.path.splitext(os.path.relpath(tif, images_root))[0]
        truth = os.path.join(truth_root, sample_id + ".json")
        if os.path.isfile(truth):
            pairs.append((tif, truth, sample_id))
        else:
            orphans += 1
            print(f"Warning: {tif} has no ground-truth JSON; skipped")
    if not pairs:
        _die("no image/JSON pairs found — check the dataset layout and the file extensions")
    print(f"Prepared {len(pairs)} image/JSON pairs for evaluation ({orphans} images skipped).")
    return pairs


def _evaluate(model, pairs, args):
    """Returns (overall MeanIoU, [(id, IoU)] of the samples below --iou_threshold)."""
    from unet_b200 import imaging
    from unet_b200.keras_api import MeanIoU
    h, w = model.spec.input_size[:2]
    overall = MeanIoU(num_classes=2, name="overall_mean_iou")      # reference :237
    below = []
    group = max(1, args.batch)
    seen = 0
    for start in range(0, len(pairs), group):
        images, truths, ids = [], [], []
        for tif, truth_json, sample_id in pairs[start:start + group]:
            seen += 1
            print(f"\r[{seen}/{len(pairs)}] {sample_id}", end="")
            try:
                image, _ = imaging.read_image_for_model(tif, h, w)
                truth = imaging.quad_mask(truth_json, h, w)
            except Exception as exc:            # unreadable image / malformed JSON: reported, not fatal (reference :244-250)
                print(f"\n{sample_id}: cannot be loaded ({exc}); skipped")
                continue
            if image is None or truth is None:
                print(f"\n{sample_id}: cannot be loaded; skipped")
                continue
            images.append(image); truths.append(truth); ids.append(sample_id)
        if not images:
            continue
        truth_batch = np.concatenate(truths, 0)
        if args.host_metrics:      # the reference's order of operations: probabilities to the host, NumPy threshold and sums
            probs = model.predict(np.concatenate(images, 0), batch_size=len(images), verbose=0)
            masks = (probs > args.pred_threshold).astype(np.uint8)   # reference :260
            scores = [imaging.sample_iou(truth_batch[k], masks[k]) for k in range(len(ids))]
            overall.update_state(truth_batch, masks)
        else:                      # same numbers from integer counts made on the device: only 32 bytes per sample come back
            prob_dev = model.predict_on_device(model._stage_in(np.concatenate(images, 0), "bench_x"))
            counts = imaging.gpu_batch_sample_counts(truth_batch, prob_dev, args.pred_threshold)
            scores = [imaging.sample_iou_from_counts(counts[k]) for k in range(len(ids))]
            overall.add_confusion(counts.sum(0))
        for sample_id, score in zip(ids, scores):
            if score < args.iou_threshold:
                below.append((sample_id, score))
                print(f"\n  IoU {score:.3f} < {args.iou_threshold:.2f}: {sample_id}")
    print()
    return float(overall.result().numpy()), below


def _report(below, args):
    if not below:
        print(f"Every sample reached the IoU threshold ({args.iou_threshold:.2f}).")
        return
    below.sort(key=lambda item: item[1])
    print(f"{len(below)} samples below the IoU threshold ({args.iou_threshold:.2f}):")
    for sample_id, score in below:
        print(f"  {score:.4f}  {sample_id}")
    if not args.low_score_log:
        return
    try:
        folder = os.path.dirname(args.low_score_log)
        if folder:
            os.makedirs(folder, exist_ok=True)
        with open(args.low_score_log, "w") as fh:
            fh.write("FileID,MeanIoU_Score\n")                     # the reference's CSV header
            fh.writelines(f"{sample_id},{score:.4f}\n" for sample_id, score in below)
        print(f"Low-score list written to {args.low_score_log}")
    except OSError as exc:
        print(f"Could not write {args.low_score_log}: {exc}")


def main(argv=None):
    args = parse_args(argv)
    started = time.time()
    images_root, truth_root = _validate(args)

    from unet_b200.keras_api import load_model
    from utils.loss import dice_loss
    from utils.metrics import dice_coef

    print(f"Loading {args.model}")
    try:
        model = load_model(args.model, custom_objects={"dice_loss": dice_loss, "dice_coef": dice_coef}, compile=False)
    except Exception as exc:
        _die(f"the model could not be loaded: {exc}")
    pairs = _collect_pairs(images_root, truth_root)
    print(f"Evaluating with prediction threshold {args.pred_threshold:.2f}")
    overall, below = _evaluate(model, pairs, args)
    bar = "=" * 30
    print(f"{bar}\nOverall Mean IoU: {overall:.4f}\n{bar}")
    _report(below, args)
    print(f"Done in {time.time() - started:.2f} s.")


if __name__ == "__main__":
    main()
