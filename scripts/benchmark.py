#!/usr/bin/env python3
"""benchmark.py — drop-in for the reference scripts/benchmark.py (accuracy evaluation) on the B200 engine.

Same CLI (reference scripts/benchmark.py:59-93): input_dir --model --iou_threshold 0.9 --pred_threshold 0.5
--low_score_log; pairs images/**/*.tif with ground_truth/**.json, builds the quad mask, thresholds the prediction,
per-sample IoU, MeanIoU(num_classes=2) over all samples, optional low-score CSV `FileID,MeanIoU_Score`.
Extension: --batch N evaluates N images per model.predict call (the reference runs batch 1; results are identical).
"""
from __future__ import annotations

import argparse
import os
import sys
import time
from glob import glob

import numpy as np

PROJECT_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if PROJECT_ROOT not in sys.path:
    sys.path.append(PROJECT_ROOT)

IMG_HEIGHT = 256
IMG_WIDTH = 256


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Benchmark a U-Net segmentation model using JSON ground truth.")
    p.add_argument("input_dir", type=str, help="Top-level directory containing 'images/' and 'ground_truth/' subfolders.")
    p.add_argument("--model", type=str, default="./models/model.h5", help="Path to the trained Keras (.h5 or .keras) model file.")
    p.add_argument("--iou_threshold", type=float, default=0.9,
                   help="Log filenames where the sample's MeanIoU is BELOW this threshold.")
    p.add_argument("--pred_threshold", type=float, default=0.5,
                   help="Threshold (0-1) to convert model's probability prediction to a binary mask for IoU calculation.")
    p.add_argument("--low_score_log", type=str, default=None,
                   help="Optional file path to save the list of files scoring below the iou_threshold.")
    p.add_argument("--batch", type=int, default=1, help="Images per predict call (extension; default 1 as the reference).")
    return p.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    t_start = time.time()
    if not os.path.isdir(args.input_dir):
        print(f"Error: Input directory not found -> {args.input_dir}")
        sys.exit(1)
    images_root = os.path.join(args.input_dir, "images")
    gtruth_root = os.path.join(args.input_dir, "ground_truth")
    if not (os.path.isdir(images_root) and os.path.isdir(gtruth_root)):
        print(f"Error: '{images_root}' or '{gtruth_root}' not found.")
        sys.exit(1)
    if not os.path.isfile(args.model):
        print(f"Error: Model file not found -> {args.model}")
        sys.exit(1)
    if not (0.0 <= args.pred_threshold <= 1.0):
        print(f"Error: Prediction threshold must be between 0.0 and 1.0 -> {args.pred_threshold}")
        sys.exit(1)
    if not (0.0 <= args.iou_threshold <= 1.0):
        print(f"Error: IoU threshold must be between 0.0 and 1.0 -> {args.iou_threshold}")
        sys.exit(1)

    from unet_b200 import imaging
    from unet_b200.keras_api import MeanIoU, load_model
    from utils.loss import dice_loss
    from utils.metrics import dice_coef

    print(f"Loading model: {args.model} ...")
    custom = {"dice_loss": dice_loss, "dice_coef": dice_coef}
    print(f"Using custom_objects for load_model: {list(custom.keys())}")
    try:
        model = load_model(args.model, custom_objects=custom, compile=False)
        print("Model loaded successfully.")
    except Exception as e:
        print("\n--- Error loading model ---")
        print(f"{e}")
        print("---------------------------\n")
        sys.exit(1)
    h, w = model.spec.input_size[:2]

    print("Finding image and ground truth pairs...")
    image_files = sorted(glob(os.path.join(images_root, "**", "*.tif"), recursive=True))
    print(f"Found {len(image_files)} '.tif' images.")
    pairs, skipped = [], 0
    for img_path in image_files:
        base = os.path.splitext(os.path.relpath(img_path, images_root))[0]
        json_path = os.path.join(gtruth_root, base + ".json")
        if os.path.isfile(json_path):
            pairs.append((img_path, json_path, base))
        else:
            print(f"Warning: No corresponding JSON found for {img_path}. Skipping.")
            skipped += 1
    if not pairs:
        print("Error: No valid image/JSON pairs found. Check dataset structure and file extensions.")
        sys.exit(1)
    print(f"Prepared {len(pairs)} image/JSON pairs for evaluation ({skipped} images skipped).")

    metric = MeanIoU(num_classes=2, name="overall_mean_iou")
    low = []
    print(f"Evaluating model (Prediction Threshold: {args.pred_threshold:.2f})...")
    bs = max(1, args.batch)
    done = 0
    for lo in range(0, len(pairs), bs):
        xs, ts, ids = [], [], []
        for img_path, json_path, fid in pairs[lo:lo + bs]:
            done += 1
            print(f"\rProcessing [{done}/{len(pairs)}]: {fid}", end="")
            try:
                x, _ = imaging.read_image_for_model(img_path, h, w)
                t = imaging.quad_mask(json_path, h, w)
            except Exception as e:
                print(f"\nError processing {fid}: {e}")
                x = t = None
            if x is None or t is None:
                print(f"\nSkipping pair due to loading error: {fid}")
                continue
            xs.append(x); ts.append(t); ids.append(fid)
        if not xs:
            continue
        prob = model.predict(np.concatenate(xs, 0), batch_size=len(xs), verbose=0)
        pred = (prob > args.pred_threshold).astype(np.uint8)
        true = np.concatenate(ts, 0)
        for k, fid in enumerate(ids):
            s = imaging.sample_iou(true[k], pred[k])
            if s < args.iou_threshold:
                low.append((fid, s))
                print(f"\nBelow threshold (IoU={s:.3f}): {fid}")
        metric.update_state(true, pred)
    print("\nEvaluation complete.")
    final = metric.result().numpy()
    print(f"\n{'=' * 30}\nOverall Mean IoU: {final:.4f}\n{'=' * 30}")
    if low:
        print(f"\nFiles scoring below IoU threshold ({args.iou_threshold:.2f}):")
        low.sort(key=lambda it: it[1])
        for fid, s in low:
            print(f"  - IoU: {s:.4f} | File: {fid}")
        if args.low_score_log:
            print(f"\nSaving low score list to: {args.low_score_log}")
            try:
                d = os.path.dirname(args.low_score_log)
                if d:
                    os.makedirs(d, exist_ok=True)
                with open(args.low_score_log, "w") as f:
                    f.write("FileID,MeanIoU_Score\n")
                    for fid, s in low:
                        f.write(f"{fid},{s:.4f}\n")
            except Exception as e:
                print(f"Error saving low score log: {e}")
    else:
        print(f"\nNo files scored below the IoU threshold ({args.iou_threshold:.2f}).")
    print(f"\nTotal benchmark time: {time.time() - t_start:.2f} seconds.")
    print("Benchmark script finished.")


if __name__ == "__main__":
    main()
