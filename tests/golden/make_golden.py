#!/usr/bin/env python3
"""Regenerates tests/golden/*.npz from the CPU oracle (oracle/unet_ref.py, float64).

PARITY UNPINNED: TensorFlow/Keras cannot be installed in this image (no wheel, no network) and the reference ships
neither tests nor weights, so these vectors are outputs of OUR restatement of the reference semantics, frozen so that
the oracle itself cannot drift silently and so that the GPU tests have a fixed target that does not depend on the
oracle code at run time.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import unet_ref as R  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(name, shape, nc, bn, rate, batch, wseed, dseed, loss="dice"):
    specs = R.layer_specs(shape, nc, rate, bn)
    P = R.init_params(specs, seed=wseed, trained_like=True)
    x, y = R.synthetic_batch(batch, shape[0], shape[1], shape[2], nc, seed=dseed)
    orc = R.UNetOracle(shape, nc, rate, bn, dtype=np.float64)
    probs = orc.forward(P, x, training=False)
    seeds = {"bneck_dropout": 11, "dec4_dropout": 12, "dec3_dropout": 13, "dec2_dropout": 14}
    lv, tprobs, grads, stats = orc.loss_and_grads(P, x, y, loss=loss, drop_seeds=seeds)
    m = R.MeanIoU(max(2, nc))
    if nc == 1:
        m.update_state(y, (probs > 0.5).astype(np.float32))
    else:
        m.update_state(y.argmax(-1), probs.argmax(-1))
    np.savez_compressed(
        os.path.join(OUT, name + ".npz"),
        shape=np.array(shape), num_classes=nc, use_batch_norm=bn, dropout_rate=rate, batch=batch, weight_seed=wseed,
        data_seed=dseed, loss_kind=loss, drop_seeds=np.array([11, 12, 13, 14]),
        probs_infer=probs.astype(np.float32), probs_train=tprobs.astype(np.float32), loss=np.float64(lv),
        dice_infer=np.float64(R.dice_coef(y, probs, dtype=np.float64)), iou_infer=np.float64(R.iou_coef(y, probs, dtype=np.float64)),
        mean_iou_infer=np.float64(m.result()),
        grad_norms=np.array([np.linalg.norm(grads[k]) for k in sorted(grads)]),
        grad_head_kernel=grads["output_mask/kernel"].astype(np.float64),
        grad_enc1_pw=grads["enc1_block1_sepconv/pointwise_kernel"].astype(np.float64),
        new_moving_mean_enc1=(stats["enc1_block1_bn/moving_mean"] if bn else np.zeros(1)))
    print(name, "loss", lv)


if __name__ == "__main__":
    case("unet_binary_32x48", (32, 48, 3), 1, True, 0.2, 2, 3, 5)
    case("unet_8class_32x32", (32, 32, 3), 8, True, 0.0, 2, 4, 6)
    case("unet_nobn_iou_32x32", (32, 32, 3), 1, False, 0.2, 2, 5, 7, loss="iou")
