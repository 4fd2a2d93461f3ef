"""The reference-facing surface on a GPU: golden vectors, Keras-shaped Model (compile / fit / predict / save / load),
utils.loss / utils.metrics, MeanIoU, and the three CLIs on synthetic directory trees in the reference's layouts."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import unet_ref as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def dev(a):
    return torch.tensor(np.asarray(a, dtype=np.float32), device="cuda")


@pytest.mark.parametrize("name", ["unet_binary_32x48", "unet_8class_32x32", "unet_nobn_iou_32x32"])
@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_engine_matches_golden(name, dtype):
    from unet_b200.engine import UNetEngine
    g = np.load(os.path.join(GOLD, name + ".npz"))
    shape = tuple(int(v) for v in g["shape"]); nc = int(g["num_classes"]); bn = bool(g["use_batch_norm"])
    rate = float(g["dropout_rate"])
    P = R.init_params(R.layer_specs(shape, nc, rate, bn), seed=int(g["weight_seed"]), trained_like=True)
    x, y = R.synthetic_batch(int(g["batch"]), shape[0], shape[1], shape[2], nc, seed=int(g["data_seed"]))
    eng = UNetEngine(shape, nc, rate, bn, dtype=dtype)
    eng.dropout_masks_from_step = False
    eng._drop_seed = dict(zip(["bneck_dropout", "dec4_dropout", "dec3_dropout", "dec2_dropout"], (int(v) for v in g["drop_seeds"])))
    eng.set_weights(P)
    probs = eng.forward_inference(dev(x)).cpu().numpy()
    tol = 1e-4 if dtype == "fp32" else 2e-2
    assert np.abs(probs - g["probs_infer"]).max() <= tol
    out3 = eng.evaluate_batch(dev(x), dev(y)).cpu().numpy()
    assert abs(out3[1] - float(g["dice_infer"])) <= 1e-3 and abs(out3[2] - float(g["iou_infer"])) <= 1e-3
    lv = eng.train_forward_backward(dev(x), dev(y), loss=str(g["loss_kind"])).cpu().numpy()[0]
    assert abs(lv - float(g["loss"])) <= (1e-4 if dtype == "fp32" else 1e-3)
    if dtype == "fp32":
        np.testing.assert_allclose(eng.wview("output_mask/kernel", eng.g).cpu().numpy().reshape(g["grad_head_kernel"].shape),
                                   g["grad_head_kernel"], rtol=5e-3, atol=2e-6)
        gp = eng.wview("enc1_block1_sepconv/pointwise_kernel", eng.g).cpu().numpy().reshape(g["grad_enc1_pw"].shape)
        assert np.linalg.norm(gp - g["grad_enc1_pw"]) <= 3e-2 * np.linalg.norm(g["grad_enc1_pw"])   # fp32 atomics order varies run to run
        if bn:
            np.testing.assert_allclose(eng.wview("enc1_block1_bn/moving_mean").cpu().numpy(), g["new_moving_mean_enc1"], rtol=1e-4, atol=1e-6)


def test_metric_and_loss_functions_match_reference_definitions():
    sys.path.insert(0, ROOT)
    from utils.loss import dice_loss, iou_loss, jaccard_loss
    from utils.metrics import dice_coef, iou_coef
    rng = np.random.default_rng(0)
    t = (rng.random((3, 20, 24, 2)) > 0.6).astype(np.float32)
    p = rng.random((3, 20, 24, 2)).astype(np.float32)
    assert float(dice_coef(t, p)) == pytest.approx(float(R.dice_coef(t, p, dtype=np.float64)), abs=1e-6)
    assert float(iou_coef(t, p)) == pytest.approx(float(R.iou_coef(t, p, dtype=np.float64)), abs=1e-6)
    assert float(dice_loss(t, p)) == pytest.approx(1 - float(R.dice_coef(t, p, dtype=np.float64)), abs=1e-6)
    assert float(iou_loss(t, p)) == pytest.approx(1 - float(R.iou_coef(t, p, dtype=np.float64)), abs=1e-6)
    assert jaccard_loss is iou_loss
    assert float(dice_coef(np.ones((1, 4, 4, 1)), np.ones((1, 4, 4, 1)))) == pytest.approx(1.0)
    assert float(dice_coef(np.zeros((1, 4, 4, 1)), np.zeros((1, 4, 4, 1)))) == pytest.approx(1.0)
    assert dice_coef(t, torch.tensor(p)).numpy().dtype == np.float32        # tensors accepted, `.numpy()` like an eager tensor
    with pytest.raises(ValueError):
        dice_coef(t, p[:, :10])


def test_mean_iou_metric_object():
    from unet_b200.keras_api import MeanIoU
    rng = np.random.default_rng(1)
    t = rng.integers(0, 2, (2, 16, 16, 1)).astype(np.uint8)
    p = rng.random((2, 16, 16, 1)).astype(np.float32); p[0, :4] = 1.0
    m, ref = MeanIoU(num_classes=2, name="mean_io_u"), R.MeanIoU(2)
    for _ in range(2):
        m.update_state(t, p); ref.update_state(t, p)              # raw probabilities: truncation semantics of train.py:231
    assert float(m.result()) == pytest.approx(ref.result(), abs=1e-7)
    np.testing.assert_array_equal(m.confusion_matrix(), ref.cm)
    m.reset_state(); ref.reset_state()
    pb = (p > 0.5).astype(np.uint8)
    m.update_state(t, pb); ref.update_state(t, pb)                 # benchmark.py:260-269 usage
    assert float(m.result().numpy()) == pytest.approx(ref.result(), abs=1e-7)


def test_model_fit_predict_save_load(tmp_path):
    sys.path.insert(0, ROOT)
    from model.u_net import U_NET
    from unet_b200.data import synthetic_batches
    from unet_b200.keras_api import (AdamW, EarlyStopping, MeanIoU, ModelCheckpoint, ReduceLROnPlateau, TensorBoard, load_model)
    from utils.loss import dice_loss
    from utils.metrics import dice_coef
    model = U_NET((64, 64, 3), num_classes=1)
    model.compile(optimizer=AdamW(learning_rate=2e-3, weight_decay=1e-4), loss=dice_loss,
                  metrics=[MeanIoU(num_classes=2, name="mean_io_u"), dice_coef])
    path = str(tmp_path / "models" / "model.h5")
    es = EarlyStopping(monitor="val_mean_io_u", patience=10, mode="max", restore_best_weights=True)
    cbs = [ModelCheckpoint(filepath=path, monitor="val_mean_io_u", mode="max", save_best_only=True, save_weights_only=False),
           es, ReduceLROnPlateau(monitor="val_mean_io_u", factor=0.2, patience=3, mode="max", min_lr=1e-6),
           TensorBoard(log_dir=str(tmp_path / "logs"), histogram_freq=1)]
    hist = model.fit(synthetic_batches(4, 64, 64, seed=1), epochs=3, steps_per_epoch=6,
                     validation_data=synthetic_batches(4, 64, 64, seed=2), validation_steps=2, callbacks=cbs, verbose=0)
    h = hist.history
    assert set(h) >= {"loss", "mean_io_u", "dice_coef", "val_loss", "val_mean_io_u", "val_dice_coef"}
    assert len(h["loss"]) == 3 and h["loss"][-1] < h["loss"][0]
    assert all(abs((1 - d) - l) < 1e-5 for d, l in zip(h["dice_coef"], h["loss"]))       # dice_loss = 1 - dice_coef
    assert os.path.isfile(path) and es.stopped_epoch == 0
    assert any(f.startswith("events.out.tfevents") or f == "scalars.csv" for f in os.listdir(tmp_path / "logs" / "train"))
    x = next(synthetic_batches(5, 64, 64, seed=3))[0]
    p1 = model.predict(x, verbose=0)
    assert p1.shape == (5, 64, 64, 1) and p1.dtype == np.float32 and 0 <= p1.min() and p1.max() <= 1
    np.testing.assert_array_equal(model.predict(x, batch_size=2, verbose=0), p1)          # batching does not change results
    for ext in (".h5", ".keras"):
        q = str(tmp_path / ("again" + ext))
        model.save(q)
        m2 = load_model(q, custom_objects={"dice_loss": dice_loss, "dice_coef": dice_coef}, compile=False)
        np.testing.assert_array_equal(m2.predict(x, verbose=0), p1)
    with pytest.raises(ValueError):
        model.predict(np.zeros((1, 32, 32, 3), np.float32))
    w = model.get_weights()
    assert len(w) == 82 + 36 and w[0].shape == (3, 3, 3, 1)
    model.get_layer("output_mask").set_weights([np.zeros((1, 1, 64, 1), np.float32), np.zeros(1, np.float32)])
    assert np.allclose(model.predict(x[:1], verbose=0), 0.5)


def _run(args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable] + args, cwd=cwd, env=env, capture_output=True, text=True, timeout=600)


def test_cli_train_inference_benchmark(tmp_path):
    import cv2
    cwd = str(tmp_path)
    # dataset in the reference's layout (scripts/train.py:79-88)
    rng = np.random.default_rng(0)
    for split, n in (("train", 6), ("val", 2)):
        fd = tmp_path / "dataset" / "train" / f"{split}_frames" / "image"
        md = tmp_path / "dataset" / "train" / f"{split}_masks" / "image"
        fd.mkdir(parents=True); md.mkdir(parents=True)
        for i in range(n):
            img = rng.integers(0, 255, (80, 100, 3), dtype=np.uint8)
            mask = np.zeros((80, 100), np.uint8); mask[20:60, 30:80] = 255
            img[20:60, 30:80] //= 4
            cv2.imwrite(str(fd / f"{i}.jpg"), img); cv2.imwrite(str(md / f"{i}.png"), mask)
    r = _run([os.path.join(ROOT, "scripts", "train.py"), "--epochs", "2", "--batch-size", "2", "--model-out", "./models/m.h5",
              "--image-size", "64"], cwd)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Steps per epoch: 3, Validation steps: 1" in r.stdout and "--- Training complete ---" in r.stdout
    assert os.path.isfile(tmp_path / "models" / "m.h5")
    # inference CLI
    img = rng.integers(0, 255, (120, 160, 3), dtype=np.uint8)
    cv2.imwrite(str(tmp_path / "in.png"), img)
    r = _run([os.path.join(ROOT, "scripts", "inference.py"), "in.png", "--model", "models/m.h5", "--output_mask", "out/mask.png",
              "--output_cropped", "out/crop.png", "--min_area", "0"], cwd)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    m = cv2.imread(str(tmp_path / "out" / "mask.png"), cv2.IMREAD_UNCHANGED)
    assert m.shape == (120, 160) and set(np.unique(m).tolist()) <= {0, 255}
    assert "Inference script finished." in r.stdout
    r = _run([os.path.join(ROOT, "scripts", "inference.py"), "in.png", "--model", "models/m.h5", "--output_mask", "out/mask_gpu.png",
              "--output_cropped", "out/crop_gpu.png", "--min_area", "0", "--gpu-prepost"], cwd)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    mg = cv2.imread(str(tmp_path / "out" / "mask_gpu.png"), cv2.IMREAD_UNCHANGED)
    assert mg.shape == m.shape and (mg == m).mean() > 0.999          # same arithmetic as the cv2 path
    r = _run([os.path.join(ROOT, "scripts", "inference.py"), "in.png", "--model", "models/m.h5", "--threshold", "1.0"], cwd)
    assert r.returncode == 1 and "Threshold must be between" in r.stdout
    # benchmark CLI (layout of scripts/prepare_dataset.py: images/**.tif + ground_truth/**.json)
    (tmp_path / "bench" / "images" / "a").mkdir(parents=True); (tmp_path / "bench" / "ground_truth" / "a").mkdir(parents=True)
    for i in range(3):
        cv2.imwrite(str(tmp_path / "bench" / "images" / "a" / f"{i}.tif"), rng.integers(0, 255, (90, 120, 3), dtype=np.uint8))
        (tmp_path / "bench" / "ground_truth" / "a" / f"{i}.json").write_text(json.dumps({"quad": [[10, 10], [100, 10], [100, 80], [10, 80]]}))
    outs = []
    for batch in ("1", "3"):
        r = _run([os.path.join(ROOT, "scripts", "benchmark.py"), "bench", "--model", "models/m.h5", "--low_score_log", "low/log.csv",
                  "--batch", batch], cwd)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        line = [l for l in r.stdout.splitlines() if l.startswith("Overall Mean IoU:")]
        assert line and "Prepared 3 image/JSON pairs" in r.stdout
        outs.append(line[0])
    assert outs[0] == outs[1]
    assert open(tmp_path / "low" / "log.csv").readline().strip() == "FileID,MeanIoU_Score"
