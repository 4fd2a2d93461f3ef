#!/usr/bin/env python3
"""train.py — drop-in for the reference scripts/train.py on the B200 engine.

Same CLI (reference scripts/train.py:92-117): --epochs 30 --batch-size 2 --learning-rate 2e-3 --weight-decay 1e-4
--model-out ./models/model.h5; same dataset layout (dataset/train/{train,val}_{frames,masks}/image), seeds, model
(U_NET((256,256,3), num_classes=1)), AdamW + dice_loss, metrics [MeanIoU(2, 'mean_io_u'), dice_coef], callbacks
(ModelCheckpoint best-only on val_mean_io_u, EarlyStopping(10, restore best), ReduceLROnPlateau(0.2, 3, 1e-6),
TensorBoard) and exit codes (1 on any error or Ctrl-C).  Extra, optional flags (defaults keep reference behaviour):
--image-size, --num-classes, --dtype {bf16,fp32}, --synthetic N (N synthetic samples instead of the dataset directories).
Multi-GPU: launch with torchrun; each rank trains on its shard of every batch, gradients are all-reduced over NCCL.
"""
from __future__ import annotations

import argparse
import os
import random as rn
import sys
import time

import numpy as np

PROJECT_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if PROJECT_ROOT not in sys.path:
    sys.path.append(PROJECT_ROOT)

DEFAULT_EPOCHS = 30
DEFAULT_BATCHSIZE = 2
DEFAULT_LR = 2e-3
DEFAULT_WEIGHT_DECAY = 1e-4
DEFAULT_MODEL_OUT = "./models/model.h5"
SEED = 2301
TRAIN_FRAMES_DIR = "dataset/train/train_frames/image"
TRAIN_MASKS_DIR = "dataset/train/train_masks/image"
VAL_FRAMES_DIR = "dataset/train/val_frames/image"
VAL_MASKS_DIR = "dataset/train/val_masks/image"
IMAGE_HEIGHT = IMAGE_WIDTH = 256
IMAGE_CHANNELS = 3
NUM_CLASSES = 1


def parse_args(argv=None) -> argparse.Namespace:
    p = argparse.ArgumentParser(description="Train a U-Net model for binary segmentation using AdamW.")
    p.add_argument("--epochs", type=int, default=DEFAULT_EPOCHS, help=f"Number of training epochs (default: {DEFAULT_EPOCHS}).")
    p.add_argument("--batch-size", type=int, default=DEFAULT_BATCHSIZE, help=f"Batch size (default: {DEFAULT_BATCHSIZE}).")
    p.add_argument("--learning-rate", type=float, default=DEFAULT_LR,
                   help=f"Initial learning rate for AdamW optimizer (default: {DEFAULT_LR}).")
    p.add_argument("--weight-decay", type=float, default=DEFAULT_WEIGHT_DECAY,
                   help=f"Weight decay for AdamW optimizer (default: {DEFAULT_WEIGHT_DECAY}).")
    p.add_argument("--model-out", type=str, default=DEFAULT_MODEL_OUT,
                   help=f"File path to save the best trained model (default: {DEFAULT_MODEL_OUT}).")
    # extensions
    p.add_argument("--image-size", type=int, default=IMAGE_HEIGHT, help="Square input size (default 256, the reference's constant).")
    p.add_argument("--num-classes", type=int, default=NUM_CLASSES, help="Output classes (default 1 = binary, sigmoid).")
    p.add_argument("--dtype", choices=["bf16", "fp32"], default="bf16", help="Activation storage type on the GPU.")
    p.add_argument("--synthetic", type=int, default=0, help="Train on N synthetic samples instead of ./dataset.")
    return p.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    os.environ["PYTHONHASHSEED"] = "0"
    np.random.seed(SEED)
    rn.seed(SEED)

    from unet_b200 import dist as D
    from unet_b200.data import PairedDirectoryIterator, list_images, synthetic_batches
    from unet_b200.keras_api import AdamW, EarlyStopping, MeanIoU, ModelCheckpoint, ReduceLROnPlateau, TensorBoard
    from model.u_net import U_NET
    from utils.loss import dice_loss
    from utils.metrics import dice_coef

    rank, local_rank, world = D.init_from_env()
    size = (args.image_size, args.image_size)
    bs = args.batch_size
    if world > 1 and (bs < world or bs % world != 0):
        # every rank must get the same, non-empty share of every batch: an empty shard cannot run a step (the other ranks
        # would wait for it in the gradient all-reduce) and unequal shards would be mis-weighted by the 1/world average
        print(f"Error: --batch-size ({bs}) must be a positive multiple of the number of GPUs ({world}) for data-parallel training.")
        sys.exit(1)
    try:
        if args.synthetic:
            n_train, n_val = args.synthetic, max(bs, args.synthetic // 5)
            lo, hi = D.shard_range(bs, rank, world)
            train_gen = synthetic_batches(hi - lo, size[0], size[1], args.num_classes, seed=SEED + rank)
            val_gen = synthetic_batches(hi - lo, size[0], size[1], args.num_classes, seed=SEED + 1000 + rank)
        else:
            for d in (TRAIN_FRAMES_DIR, TRAIN_MASKS_DIR, VAL_FRAMES_DIR, VAL_MASKS_DIR):
                if not os.path.isdir(d):
                    raise FileNotFoundError(f"Directory not found: {d}")
            print("Setting up data generators...")
            train_it = PairedDirectoryIterator(TRAIN_FRAMES_DIR, TRAIN_MASKS_DIR, size, bs, shuffle=True,
                                               horizontal_flip=True, seed=SEED)
            val_it = PairedDirectoryIterator(VAL_FRAMES_DIR, VAL_MASKS_DIR, size, bs, shuffle=False, seed=SEED)
            n_train, n_val = train_it.samples, val_it.samples
            lo, hi = D.shard_range(bs, rank, world)
            # Keras' iterator yields the short last batch of each pass; data parallel only takes full batches (see above),
            # and every rank skips the same ones, so the ranks stay in lock-step
            full = (lambda it: it) if world == 1 else (lambda it: (b for b in it if len(b[0]) == bs))
            train_gen = ((x[lo:hi], y[lo:hi]) for x, y in full(train_it))
            val_gen = ((x[lo:hi], y[lo:hi]) for x, y in full(val_it))
    except Exception as e:
        print("\n--- Error setting up data generators ---")
        print(f"{e}")
        sys.exit(1)
    if n_train == 0:
        print(f"Error: No training images found in {TRAIN_FRAMES_DIR}")
        sys.exit(1)
    if n_val == 0:
        print(f"Error: No validation images found in {VAL_FRAMES_DIR}")
        sys.exit(1)
    if world > 1 and (n_train < bs or n_val < bs):
        print(f"Error: data-parallel training needs at least one full batch ({bs}) of training ({n_train}) and validation ({n_val}) images.")
        sys.exit(1)

    print("Building U-Net model...")
    model = U_NET((size[0], size[1], IMAGE_CHANNELS), num_classes=args.num_classes)
    model.dtype_name = args.dtype
    print(f"Compiling model with AdamW (LR={args.learning_rate}, WD={args.weight_decay})...")
    model.compile(optimizer=AdamW(learning_rate=args.learning_rate, weight_decay=args.weight_decay), loss=dice_loss,
                  metrics=[MeanIoU(num_classes=max(2, args.num_classes), name="mean_io_u"), dice_coef])
    if rank == 0:
        model.summary(line_length=100)
    if world > 1:
        import torch.distributed as dist
        dist.broadcast(model.engine.w, src=0)
        dist.broadcast(model.engine.state, src=0)
        model.engine._stage_dirty = True
        model.enable_data_parallel()

    steps_per_epoch = max(1, n_train // bs)
    validation_steps = max(1, n_val // bs)
    print(f"Steps per epoch: {steps_per_epoch}, Validation steps: {validation_steps}")
    if n_train < bs:
        print(f"Warning: Training dataset size ({n_train}) < batch size ({bs}).")
    if n_val < bs:
        print(f"Warning: Validation dataset size ({n_val}) < batch size ({bs}).")

    monitor_metric = "val_mean_io_u"
    monitor_mode = "max" if "loss" not in monitor_metric else "min"
    print(f"Setting up Callbacks - Monitoring: '{monitor_metric}' (mode: {monitor_mode})")
    callbacks = []
    early_stopping = EarlyStopping(monitor=monitor_metric, patience=10, mode=monitor_mode, restore_best_weights=True, verbose=1)
    if rank == 0:
        model_dir = os.path.dirname(args.model_out)
        if model_dir:
            os.makedirs(model_dir, exist_ok=True)
            print(f"Ensured model save directory exists: {model_dir}")
        callbacks.append(ModelCheckpoint(filepath=args.model_out, monitor=monitor_metric, mode=monitor_mode,
                                         save_best_only=True, save_weights_only=False, verbose=1))
    callbacks.append(early_stopping)
    callbacks.append(ReduceLROnPlateau(monitor=monitor_metric, factor=0.2, patience=3, mode=monitor_mode, min_lr=1e-6, verbose=1))
    if rank == 0:
        log_dir = os.path.join("./logs", time.strftime("%Y%m%d_%H%M%S"))
        os.makedirs(log_dir, exist_ok=True)
        print(f"TensorBoard logs will be saved to: {log_dir}")
        callbacks.append(TensorBoard(log_dir=log_dir, histogram_freq=1))

    print(f"\n--- Starting Training ({args.epochs} epochs) ---")
    try:
        history = model.fit(train_gen, epochs=args.epochs, steps_per_epoch=steps_per_epoch, validation_data=val_gen,
                            validation_steps=validation_steps, callbacks=callbacks, verbose=1 if rank == 0 else 0)
        print("\n--- Training complete ---")
        scores = history.history.get(monitor_metric)
        if early_stopping.stopped_epoch > 0:
            stopped = early_stopping.stopped_epoch + 1
            print(f"Early stopping triggered at epoch {stopped}")
            best, best_epoch = early_stopping.best, stopped - early_stopping.patience
        elif scores:
            best = max(scores) if monitor_mode == "max" else min(scores)
            best_epoch = int(np.argmax(scores) if monitor_mode == "max" else np.argmin(scores)) + 1
        else:
            best, best_epoch = float("nan"), "N/A"
        print(f"Best monitored score ({monitor_metric}): {best:.4f} (from epoch {best_epoch})")
        print(f"Best model saved to: {args.model_out}")
        if world > 1:
            sys.stdout.flush()
            if not D.shutdown(model.engine):
                os._exit(0)
    except KeyboardInterrupt:
        print("\n--- Training interrupted by user ---")
        print(f"Model state might not be saved correctly to {args.model_out} unless a checkpoint occurred.")
        sys.exit(1)
    except Exception as e:
        print("\n--- Error during model training ---")
        print(f"{e}")
        import traceback
        traceback.print_exc()
        print("-----------------------------------\n")
        sys.exit(1)


if __name__ == "__main__":
    main()
