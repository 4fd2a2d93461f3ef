#!/usr/bin/env python3
"""inference.py — drop-in for the reference scripts/inference.py on the B200 engine.

Same CLI (reference scripts/inference.py:54-96): input --output_mask --output_cropped --model --threshold --min_area;
same pipeline: BGR read, /255, bilinear resize to 256x256, model.predict, bilinear resize of the probability mask back
to the original size, threshold -> {0,255} PNG, largest external contour -> bounding-box crop; same exit codes.
"""
from __future__ import annotations

import argparse
import os
import sys

PROJECT_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if PROJECT_ROOT not in sys.path:
    sys.path.append(PROJECT_ROOT)

IMG_HEIGHT = 256
IMG_WIDTH = 256
MIN_CONTOUR_AREA = 100


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Perform segmentation and cropping using a trained U-Net model.")
    p.add_argument("input", type=str, help="Path to the input image file.")
    p.add_argument("--output_mask", type=str, default="./outputs_test/output_mask.png",
                   help="Output path for the predicted binary mask image (0 or 255).")
    p.add_argument("--output_cropped", type=str, default="./outputs_test/output_cropped.png",
                   help="Output path for the cropped image based on the largest mask contour.")
    p.add_argument("--model", type=str, default="./models/model.h5", help="Path to the trained Keras (.h5 or .keras) model file.")
    p.add_argument("--threshold", type=float, default=0.5,
                   help="Threshold value (0.0 to 1.0) to convert probability mask to binary mask.")
    p.add_argument("--min_area", type=float, default=MIN_CONTOUR_AREA,
                   help=f"Minimum contour area threshold for cropping (default: {MIN_CONTOUR_AREA}).")
    p.add_argument("--gpu-prepost", action="store_true",
                   help="Extension: resize/normalise and upsample/threshold on the GPU (same arithmetic as the cv2 path).")
    return p.parse_args(argv)


def _write(path: str, img, what: str) -> None:
    import cv2
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    try:
        if not cv2.imwrite(path, img):
            print(f"Warning: cv2.imwrite failed to save {what} to {path}")
    except Exception as e:
        print(f"Error saving {what}: {e}")


def postprocess_and_save_results(prob_mask_pred, original_bgr, orig_height: int, orig_width: int, output_mask_path: str,
                                 output_cropped_path: str, binary_threshold: float = 0.5,
                                 min_contour_area: float = 100.0) -> None:
    """Same name, arguments and printed messages as the reference's function (scripts/inference.py:127-201): bilinear
    resize of the (h, w, 1) probabilities to the original size, threshold -> {0,255} mask PNG, largest external contour ->
    bounding-box crop of the original BGR image.  `prob_mask_pred` may be a host array (cv2 resize, as the reference) or a
    CUDA tensor (the --gpu-prepost path: unet_postprocess_mask does the resize + threshold on the device)."""
    from unet_b200 import imaging
    if prob_mask_pred is None or original_bgr is None:
        print("Error: Invalid input provided for postprocessing.")
        return
    print("Processing predicted mask...")
    if hasattr(prob_mask_pred, "is_cuda") and prob_mask_pred.is_cuda:
        mask = imaging.gpu_probability_to_mask(prob_mask_pred, orig_height, orig_width, binary_threshold)
    else:
        mask = imaging.probability_to_mask(prob_mask_pred, orig_height, orig_width, binary_threshold)
    print(f"Saving binary mask to {output_mask_path} ...")
    _write(output_mask_path, mask, "mask")
    print("Finding largest contour for cropping...")
    crop, area, rect = imaging.largest_region_crop(mask, original_bgr, min_contour_area)
    if area is None:
        print("No contours found in the binary mask. Cropped image not saved.")
    elif crop is None:
        print(f"Largest contour area ({area:.0f}) is below minimum threshold ({min_contour_area:.0f}). Cropped image not saved.")
    else:
        x0, y0, cw, ch = rect
        print(f"Largest contour area: {area:.0f} > {min_contour_area:.0f}. Cropping region: (x={x0}, y={y0}, w={cw}, h={ch})")
        print(f"Saving cropped image to {output_cropped_path} ...")
        _write(output_cropped_path, crop, "cropped image")


def main(argv=None):
    args = parse_args(argv)
    if not os.path.isfile(args.input):
        print(f"Error: Input image not found -> {args.input}")
        sys.exit(1)
    if not os.path.isfile(args.model):
        print(f"Error: Model file not found -> {args.model}")
        sys.exit(1)
    if not (0.0 < args.threshold < 1.0):
        print(f"Error: Threshold must be between 0.0 and 1.0 -> {args.threshold}")
        sys.exit(1)

    from unet_b200 import imaging
    from unet_b200.keras_api import load_model
    from utils.loss import dice_loss
    from utils.metrics import dice_coef

    print(f"Loading model from {args.model} ...")
    custom = {"dice_loss": dice_loss, "dice_coef": dice_coef}
    print(f"Using custom_objects for load_model: {list(custom.keys())}")
    try:
        model = load_model(args.model, custom_objects=custom, compile=False)
        print("Model loaded successfully.")
    except Exception as e:
        print("\n--- Error loading model ---")
        print(f"{e}")
        print(f"\nTroubleshooting:\n1. Is the model path correct? -> {args.model}")
        print("2. Was the file written by Keras (legacy .h5 / .keras) or by this package for the reference U-Net?")
        print("---------------------------\n")
        sys.exit(1)

    h, w = model.spec.input_size[:2]
    print(f"Loading and preprocessing image: {args.input} ...")
    if args.gpu_prepost:
        import cv2
        bgr = cv2.imread(args.input, cv2.IMREAD_COLOR)
        x = None if bgr is None else imaging.gpu_preprocess([bgr], h, w)
    else:
        x, bgr = imaging.read_image_for_model(args.input, h, w)
    if x is None:
        print(f"Error: Could not read image from {args.input}")
        sys.exit(1)
    print("Running prediction...")
    try:
        pred = model.predict_on_device(x) if args.gpu_prepost else model.predict(x, verbose=0)
    except Exception as e:
        print(f"Error during model prediction: {e}")
        sys.exit(1)
    if pred is None or pred.ndim != 4 or pred.shape[0] != 1:
        print(f"Error: Unexpected model prediction shape: {None if pred is None else pred.shape}")
        sys.exit(1)

    print("Postprocessing results...")
    postprocess_and_save_results(pred[0], bgr, bgr.shape[0], bgr.shape[1], args.output_mask, args.output_cropped,
                                 binary_threshold=args.threshold, min_contour_area=args.min_area)
    print("Inference script finished.")


if __name__ == "__main__":
    main()
