"""Host-side logic that needs no GPU: topology/spec, Keras-shaped model bookkeeping and summary, callbacks, model files
(HDF5 subset, .keras, .npz), the input pipeline, CPU pre/post-processing, CLI argument surfaces."""
import io
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spec_matches_reference_topology():
    from unet_b200.spec import UNetSpec
    sp = UNetSpec((256, 256, 3))
    assert len(sp.layers) == 72
    assert (sp.trainable_params, sp.non_trainable_params) == (5_988_252, 11_776)
    names = [l.name for l in sp.layers]
    assert names[0] == "input_image" and names[-1] == "output_mask"
    assert names[1:4] == ["enc1_block1_sepconv", "enc1_block1_bn", "enc1_block1_relu"]
    assert "dec1_dropout" not in names and "dec2_dropout" in names and "bneck_dropout" in names
    assert [l.kind for l in sp.layers if l.name.startswith("dec4_")][:3] == ["Conv2DTranspose", "Concatenate", "Dropout"]
    assert sp.layers[names.index("dec4_concat")].out_shape == (None, 32, 32, 1024)
    assert sp.params["dec4_upsample/kernel"].shape == (2, 2, 512, 1024)
    assert sp.layer_weight_names("enc1_block1_bn") == [f"enc1_block1_bn/{k}" for k in ("gamma", "beta", "moving_mean", "moving_variance")]
    offs = [p.offset for p in sp.params.values() if p.trainable]
    assert offs == sorted(offs) and all(o % 8 == 0 for o in offs)
    with pytest.raises(ValueError):
        UNetSpec((256, 256))


def test_u_net_builder_prints_and_validates(capsys):
    sys.path.insert(0, ROOT)
    from model.u_net import U_NET, conv_block
    with pytest.raises(ValueError, match="input_size must be a tuple"):
        U_NET((256, 256))
    m = U_NET((256, 256, 3))
    out = capsys.readouterr().out
    assert "Building Encoder..." in out and "  Decoder Stage 1, Filters: 64" in out and "U-Net model built successfully." in out
    assert m.name == "U-NET-Segmentation" and m.count_params() == 6_000_028 and len(m.layers) == 72
    assert m.output_shape == (None, 256, 256, 1)
    assert m.get_layer("output_mask").count_params() == 65
    with pytest.raises(ValueError):
        m.get_layer("nope")
    lines = []
    m.summary(line_length=100, print_fn=lines.append)
    text = "\n".join(lines)
    assert 'Model: "U-NET-Segmentation"' in text and "Total params: 6,000,028" in text
    assert "Trainable params: 5,988,252" in text and "Non-trainable params: 11,776" in text
    assert conv_block(None, 64, name_prefix="a") == [("a_sepconv", "SeparableConv2D"), ("a_bn", "BatchNormalization"), ("a_relu", "Activation")]


def test_compile_validation():
    from unet_b200.keras_api import AdamW, MeanIoU, Model

    def dice_loss(a, b): ...
    def dice_coef(a, b): ...
    def weird(a, b): ...
    m = Model((32, 32, 3))
    m.compile(optimizer=AdamW(learning_rate=2e-3, weight_decay=1e-4), loss=dice_loss, metrics=[MeanIoU(num_classes=2, name="mean_io_u"), dice_coef])
    assert m._metric_names() == ["mean_io_u", "dice_coef"] and m._loss_kind == "dice"
    with pytest.raises(ValueError):
        m.compile(optimizer=AdamW(), loss=weird)
    with pytest.raises(ValueError):
        m.compile(optimizer=AdamW(), loss=dice_loss, metrics=[weird])
    with pytest.raises(RuntimeError):
        Model((32, 32, 3)).fit(iter([]), steps_per_epoch=1)


class _FakeModel:
    def __init__(self):
        from unet_b200.keras_api import AdamW
        self.optimizer = AdamW(learning_rate=1.0)
        self.stop_training = False
        self.saved = []
        self.w = 0

    def save(self, path):
        self.saved.append(path)

    save_weights = save

    def get_weights(self):
        return self.w

    def set_weights(self, w):
        self.w = w

    def _push_hyper(self):
        pass


def test_callbacks_follow_keras_rules(tmp_path):
    from unet_b200.keras_api import EarlyStopping, ModelCheckpoint, ReduceLROnPlateau
    m = _FakeModel()
    ck = ModelCheckpoint(str(tmp_path / "m.h5"), monitor="val_mean_io_u", mode="max", save_best_only=True)
    es = EarlyStopping(monitor="val_mean_io_u", patience=2, mode="max", restore_best_weights=True)
    rl = ReduceLROnPlateau(monitor="val_mean_io_u", factor=0.2, patience=2, mode="max", min_lr=0.01)
    for cb in (ck, es, rl):
        cb.set_model(m); cb.on_train_begin()
    scores = [0.5, 0.6, 0.55, 0.58, 0.4]
    for ep, s in enumerate(scores):
        m.w = ep
        logs = {"val_mean_io_u": s}
        for cb in (ck, es, rl):
            cb.on_epoch_end(ep, logs)
        if m.stop_training:
            break
    assert len(m.saved) == 2                      # epochs 0 and 1 improved
    assert m.stop_training and es.stopped_epoch == 3 and es.best == 0.6
    es.on_train_end()
    assert m.w == 1                               # best weights restored
    assert m.optimizer.learning_rate == pytest.approx(0.2)   # one reduction after 2 epochs without improvement
    assert "learning_rate" in logs


def test_tensorboard_scalars_and_histograms(tmp_path):
    """TensorBoard(log_dir, histogram_freq=1) as scripts/train.py:299-302 builds it: scalar events for train / validation runs and
    one weight histogram per variable per epoch, readable by tensorboard's own event reader."""
    pytest.importorskip("tensorboard")
    from tensorboard.backend.event_processing.event_file_loader import EventFileLoader
    from unet_b200.keras_api import TensorBoard

    class M:
        def get_weights_dict(self):
            return {"enc1_block1_bn/gamma": np.ones(64, np.float32), "output_mask/kernel": np.linspace(-1, 1, 64, dtype=np.float32).reshape(1, 1, 64, 1)}
    tb = TensorBoard(log_dir=str(tmp_path), histogram_freq=1)
    tb.set_model(M())
    for ep in range(2):
        tb.on_epoch_end(ep, {"loss": 0.5 - 0.1 * ep, "val_loss": 0.6, "val_mean_io_u": 0.4 + 0.1 * ep})
    tb.on_train_end()
    tags = {"train": [], "validation": []}
    for run in tags:
        files = [f for f in os.listdir(tmp_path / run) if "tfevents" in f]
        assert files
        for ev in EventFileLoader(str(tmp_path / run / files[0])).Load():
            for v in ev.summary.value:
                tags[run].append((ev.step, v.tag, v.WhichOneof("value")))
    assert (0, "epoch_loss", "simple_value") in [(s, t, k) for s, t, k in tags["train"]] or any(t == "epoch_loss" for _, t, _ in tags["train"])
    assert any(t == "epoch_mean_io_u" for _, t, _ in tags["validation"])
    histos = [(s, t) for s, t, k in tags["train"] if t in ("enc1_block1_bn/gamma", "output_mask/kernel")]
    assert sorted(histos) == [(0, "enc1_block1_bn/gamma"), (0, "output_mask/kernel"), (1, "enc1_block1_bn/gamma"), (1, "output_mask/kernel")]


def test_h5lite_roundtrip_and_layout():
    from unet_b200 import h5lite
    root = h5lite.Group()
    root.attrs["model_config"] = json.dumps({"a": 1}).encode()
    root.attrs["names"] = [b"alpha", b"be"]
    g = root.group("model_weights").group("layer")
    g.attrs["weight_names"] = [b"layer/kernel:0"]
    g.dataset("kernel:0", np.arange(24, dtype=np.float32).reshape(2, 3, 4), attrs={"note": b"x"})
    root.group("many")
    for i in range(150):                          # more than one symbol-table node
        root["many"].dataset(f"d{i:03d}", np.full((3,), i, np.float64))
    root.dataset("scalar", np.float32(2.5))
    root.dataset("ints", np.arange(5, dtype=np.int64))
    buf = io.BytesIO()
    h5lite.write(buf, root)
    raw = buf.getvalue()
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0
    assert int.from_bytes(raw[40:48], "little") == len(raw)       # end-of-file address in the superblock
    back = h5lite.read(raw)
    assert json.loads(back.attrs["model_config"]) == {"a": 1}
    assert [bytes(b) for b in back.attrs["names"]] == [b"alpha", b"be"]
    k = back["model_weights/layer/kernel:0"]
    np.testing.assert_array_equal(np.asarray(k), np.arange(24, dtype=np.float32).reshape(2, 3, 4))
    assert k.attrs["note"] == b"x"
    assert len(back["many"].datasets) == 150 and float(np.asarray(back["many/d149"])[0]) == 149.0
    assert float(np.asarray(back["scalar"])) == 2.5 and np.asarray(back["ints"]).tolist() == [0, 1, 2, 3, 4]
    with pytest.raises(ValueError):
        h5lite.read(b"not an hdf5 file at all")


@pytest.mark.parametrize("ext", [".h5", ".keras", ".npz"])
@pytest.mark.parametrize("cfg", [(1, True, 0.2), (8, False, 0.0)])
def test_model_files_roundtrip(tmp_path, ext, cfg):
    from unet_b200 import weights_io
    from unet_b200.spec import UNetSpec
    nc, bn, rate = cfg
    sp = UNetSpec((32, 48, 3), nc, rate, bn)
    rng = np.random.default_rng(0)
    W = {n: rng.standard_normal(p.shape).astype(np.float32) for n, p in sp.params.items()}
    path = str(tmp_path / ("model" + ext))
    weights_io.write_model_file(path, sp, W)
    cfg_read, w = weights_io._read_any(path)
    sp2 = weights_io.spec_from_config(cfg_read)
    assert (sp2.input_size, sp2.num_classes, sp2.use_batch_norm, sp2.dropout_rate) == ((32, 48, 3), nc, bn, rate)
    for n in W:
        np.testing.assert_array_equal(w[n], W[n])
    np.testing.assert_array_equal(weights_io.read_weights(path, sp)["output_mask/kernel"], W["output_mask/kernel"])
    m = weights_io.load_model(path)              # builds the model object; weights stay pending until a GPU is used
    assert m.spec.num_classes == nc and set(m._pending_weights) == set(W)
    with pytest.raises(OSError):
        weights_io.load_model(str(tmp_path / "missing.h5"))


def test_keras_layout_of_legacy_h5(tmp_path):
    """The .h5 we write has the attribute / group structure Keras' legacy saver produces."""
    from unet_b200 import h5lite, weights_io
    from unet_b200.spec import UNetSpec
    sp = UNetSpec((32, 32, 3))
    W = {n: np.zeros(p.shape, np.float32) for n, p in sp.params.items()}
    path = str(tmp_path / "m.h5")
    weights_io.write_model_file(path, sp, W)
    root = h5lite.read(open(path, "rb").read())
    assert set(root.attrs) >= {"model_config", "keras_version", "backend"}
    mw = root["model_weights"]
    names = [bytes(b).decode() for b in mw.attrs["layer_names"]]
    assert names == [l.name for l in sp.layers]
    wn = [bytes(b).decode() for b in mw["enc1_block1_sepconv"].attrs["weight_names"]]
    assert wn == ["enc1_block1_sepconv/depthwise_kernel:0", "enc1_block1_sepconv/pointwise_kernel:0"]
    assert np.asarray(mw["enc1_block1_sepconv/enc1_block1_sepconv/depthwise_kernel:0"]).shape == (3, 3, 3, 1)
    assert len(mw["enc1_pool"].attrs["weight_names"]) == 0


def test_paired_directory_iterator(tmp_path):
    import cv2
    from unet_b200.data import PairedDirectoryIterator, synthetic_batches
    fd, md = tmp_path / "frames" / "image", tmp_path / "masks" / "image"
    fd.mkdir(parents=True); md.mkdir(parents=True)
    rng = np.random.default_rng(0)
    for i in range(5):
        img = rng.integers(0, 255, (20, 30, 3), dtype=np.uint8)
        mask = np.zeros((20, 30), np.uint8); mask[:, : 5 + 3 * i] = 255
        cv2.imwrite(str(fd / f"{i}.png"), img); cv2.imwrite(str(md / f"{i}.png"), mask)
    it = PairedDirectoryIterator(str(fd), str(md), target_size=(16, 16), batch_size=2, shuffle=True, horizontal_flip=True, seed=2301)
    gen = iter(it)
    sizes = [next(gen)[0].shape[0] for _ in range(3)]
    assert sizes == [2, 2, 1]                    # the last partial batch of an epoch is smaller, as in Keras
    x, y = next(gen)
    assert x.shape == (2, 16, 16, 3) and y.shape == (2, 16, 16, 1) and x.dtype == np.float32
    assert 0.0 <= x.min() and x.max() <= 1.0 and set(np.unique(y).tolist()) <= {0.0, 1.0}
    a = [b[1].sum() for b, _ in zip(iter(PairedDirectoryIterator(str(fd), str(md), (16, 16), 2, seed=1)), range(3))]
    b = [b[1].sum() for b, _ in zip(iter(PairedDirectoryIterator(str(fd), str(md), (16, 16), 2, seed=1)), range(3))]
    assert a == b
    xs, ys = next(synthetic_batches(3, 32, 32, classes=8))
    assert xs.shape == (3, 32, 32, 3) and ys.shape == (3, 32, 32, 8) and np.all(ys.sum(-1) == 1)
    # threaded decode + background prefetch yields the very same stream as the synchronous path
    kw = dict(target_size=(16, 16), batch_size=2, shuffle=True, horizontal_flip=True, seed=7)
    sync = [b for b, _ in zip(iter(PairedDirectoryIterator(str(fd), str(md), workers=0, **kw)), range(7))]
    gen_t = iter(PairedDirectoryIterator(str(fd), str(md), workers=4, prefetch=3, **kw))
    thr = [next(gen_t) for _ in range(7)]
    gen_t.close()                                # stops the producer thread and the pool
    for (xa, ya), (xb, yb) in zip(sync, thr):
        assert np.array_equal(xa, xb) and np.array_equal(ya, yb)
    # same values as the plain recipe: float32(image) * float32(1/255)
    first = cv2.cvtColor(cv2.imread(str(fd / "0.png")), cv2.COLOR_BGR2RGB)
    ref0 = cv2.resize(first, (16, 16), interpolation=cv2.INTER_LINEAR).astype(np.float32) * np.float32(1.0 / 255.0)
    x0, _ = next(iter(PairedDirectoryIterator(str(fd), str(md), (16, 16), 1, shuffle=False, workers=2)))
    assert np.array_equal(x0[0], ref0)
    # a decode error surfaces in the consumer
    (md / "4.png").write_bytes(b"not an image")
    with pytest.raises(OSError):
        for _ in zip(iter(PairedDirectoryIterator(str(fd), str(md), (16, 16), 2, shuffle=False, workers=2)), range(5)):
            pass


def test_postprocessing_known_answer(tmp_path):
    """Crop = bounding box of the largest external contour, cut from the ORIGINAL image (inference.py:173-187)."""
    from unet_b200 import imaging
    img = np.arange(60 * 80 * 3, dtype=np.uint32).reshape(60, 80, 3).astype(np.uint8)
    prob = np.zeros((30, 40, 1), np.float32)
    prob[10:20, 5:25] = 0.9                      # big blob
    prob[2:4, 30:33] = 0.9                       # small blob
    mask = imaging.probability_to_mask(prob, 60, 80, 0.5)
    assert mask.shape == (60, 80) and set(np.unique(mask).tolist()) == {0, 255}
    crop, area, rect = imaging.largest_region_crop(mask, img, 100)
    x, y, w, h = rect
    assert abs(x - 10) <= 1 and abs(y - 20) <= 1 and abs(w - 40) <= 2 and abs(h - 20) <= 2
    np.testing.assert_array_equal(crop, img[y:y + h, x:x + w])
    assert imaging.largest_region_crop(np.zeros((8, 8), np.uint8), img, 100) == (None, None, None)
    c2, a2, r2 = imaging.largest_region_crop(mask, img, 1e9)
    assert c2 is None and a2 == area
    assert imaging.sample_iou(np.ones((4, 4, 1)), np.zeros((4, 4))) == pytest.approx(1e-7 / (16 + 1e-7), rel=1e-2)


def test_chile_id_card_known_answer(tmp_path):
    """The one byte-exact fixture the reference holds: samples/test_images/chile_id_card.png with
    samples/usage/chile_id_card/output_{mask,cropped}.png — the reference's own post-processing output
    (scripts/inference.py:173-187).  Copied unmodified into tests/golden/chile_id_card/."""
    import cv2
    from unet_b200 import imaging
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import importlib
    inference = importlib.import_module("inference")
    gold = os.path.join(ROOT, "tests", "golden", "chile_id_card")
    bgr = cv2.imread(os.path.join(gold, "input.png"), cv2.IMREAD_COLOR)
    mask = cv2.imread(os.path.join(gold, "output_mask.png"), cv2.IMREAD_GRAYSCALE)
    want = cv2.imread(os.path.join(gold, "output_cropped.png"), cv2.IMREAD_COLOR)
    assert bgr.shape == (960, 540, 3) and mask.shape == (960, 540) and set(np.unique(mask).tolist()) == {0, 255}
    crop, area, rect = imaging.largest_region_crop(mask, bgr, 100.0)
    assert rect == (38, 296, 466, 300) and area == 128625.0
    np.testing.assert_array_equal(crop, want)                      # byte-exact
    # ... and through the script's own post-processing entry point (same name and arguments as the reference's), from a
    # probability map: one at the original size (identity resize) and one at the model's 256x256 that upsamples to the mask
    prob_full = (mask.astype(np.float32) / 255.0)[..., None]
    out_m, out_c = str(tmp_path / "m.png"), str(tmp_path / "c.png")
    inference.postprocess_and_save_results(prob_full, bgr, 960, 540, out_m, out_c, binary_threshold=0.5, min_contour_area=100.0)
    np.testing.assert_array_equal(cv2.imread(out_m, cv2.IMREAD_GRAYSCALE), mask)
    np.testing.assert_array_equal(cv2.imread(out_c, cv2.IMREAD_COLOR), want)


def test_quad_mask(tmp_path):
    import cv2
    from unet_b200 import imaging
    (tmp_path / "images").mkdir(); (tmp_path / "ground_truth").mkdir()
    cv2.imwrite(str(tmp_path / "images" / "a.tif"), np.zeros((100, 200, 3), np.uint8))
    jp = tmp_path / "ground_truth" / "a.json"
    jp.write_text(json.dumps({"quad": [[50, 25], [150, 25], [150, 75], [50, 75]]}))
    m = imaging.quad_mask(str(jp), 64, 64)
    assert m.shape == (1, 64, 64, 1) and m.dtype == np.uint8 and set(np.unique(m).tolist()) == {0, 1}
    assert abs(m.mean() - 0.25) < 0.03 and m[0, 32, 32, 0] == 1 and m[0, 2, 2, 0] == 0


def test_cli_surfaces():
    sys.path.insert(0, ROOT)
    from scripts import benchmark, inference, train
    a = train.parse_args([])
    assert (a.epochs, a.batch_size, a.learning_rate, a.weight_decay, a.model_out) == (30, 2, 2e-3, 1e-4, "./models/model.h5")
    assert train.SEED == 2301 and train.TRAIN_FRAMES_DIR == "dataset/train/train_frames/image"
    i = inference.parse_args(["img.png"])
    assert (i.output_mask, i.output_cropped, i.model, i.threshold, i.min_area) == (
        "./outputs_test/output_mask.png", "./outputs_test/output_cropped.png", "./models/model.h5", 0.5, 100)
    b = benchmark.parse_args(["data"])
    assert (b.model, b.iou_threshold, b.pred_threshold, b.low_score_log) == ("./models/model.h5", 0.9, 0.5, None)
    for mod, argv in ((inference, ["/nonexistent.png"]), (benchmark, ["/nonexistent_dir"])):
        with pytest.raises(SystemExit) as e:
            mod.main(argv)
        assert e.value.code == 1


def test_shard_range_covers_everything():
    from unet_b200.dist import shard_range
    for n in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_host_copy_paths():
    """predict()'s staging copy: NumPy path below 8 MB, torch (this process's share of the CPU threads, also when torchrun pinned
    OMP_NUM_THREADS=1) above, dtype conversion, tensor/array mixes"""
    import torch
    from unet_b200.keras_api import _host_copy
    rng = np.random.default_rng(3)
    small = rng.integers(0, 255, (2, 16, 16, 3)).astype(np.uint8)
    dst = torch.empty(small.shape, dtype=torch.float32)
    _host_copy(dst, small)
    assert np.array_equal(dst.numpy(), small.astype(np.float32))
    big = rng.random((3, 512, 512, 3)).astype(np.float64)            # 18.9 MB: the threaded path, with a down-cast
    dstb = torch.empty(big.shape, dtype=torch.float32)
    _host_copy(dstb, big)
    assert np.array_equal(dstb.numpy(), big.astype(np.float32))
    out = np.zeros((4, 512, 512, 2), np.float32)
    src = torch.arange(2 * 512 * 512 * 2, dtype=torch.float32).view(2, 512, 512, 2)
    _host_copy(out[1:3], src)                                          # into a slice of the result array
    assert np.array_equal(out[1:3], src.numpy()) and out[0].sum() == 0 and out[3].sum() == 0
    ro = rng.random((3, 512, 512, 3)).astype(np.float32); ro.setflags(write=False)
    _host_copy(dstb, ro)                                               # read-only source array
    assert np.array_equal(dstb.numpy(), ro)


def test_cast_transpose_table_layout():
    """host side of unet_cast_transpose_bf16_batched: tile offsets accumulate per matrix, destinations are validated
    (no kernel runs here: CPU tensors stand in for the device buffers)"""
    import torch
    from unet_b200 import ops
    base = torch.zeros(64 * 64 + 3 * 64 + 70 * 33)
    a = torch.zeros((64, 64), dtype=torch.bfloat16); at = torch.zeros((64, 64), dtype=torch.bfloat16)
    b_t = torch.zeros((64, 3), dtype=torch.bfloat16)
    c = torch.zeros((70, 33), dtype=torch.bfloat16)
    table, n, tiles = ops.cast_transpose_table(base, [(0, a, at, 64, 64), (4096, None, b_t, 3, 64), (4096 + 192, c, None, 70, 33)])
    assert n == 3 and tiles == 4 + 2 + 3 * 2
    assert table.dtype == torch.int64 and tuple(table.shape) == (3, 6)
    rows = table.tolist()
    assert [r[5] for r in rows] == [0, 4, 6] and [r[0] for r in rows] == [0, 4096, 4288]
    assert rows[1][1] == 0 and rows[2][2] == 0 and rows[0][1] == a.data_ptr() and rows[1][2] == b_t.data_ptr()
    with pytest.raises(ValueError):
        ops.cast_transpose_table(base, [(0, torch.zeros((64, 64)), None, 64, 64)])            # not bf16
    with pytest.raises(ValueError):
        ops.cast_transpose_table(base, [(0, None, torch.zeros((64, 3), dtype=torch.bfloat16), 64, 3)])   # dst_t must be [C, R]
    with pytest.raises(ValueError):
        ops.cast_transpose_table(base, [(base.numel() - 10, a, None, 64, 64)])                # outside the source buffer
