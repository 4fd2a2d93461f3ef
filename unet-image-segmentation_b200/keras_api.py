"""Keras-shaped host surface over the B200 engine: the subset of `tf.keras` the reference scripts touch.

Reference call sites this mirrors (same names, argument meaning and error behaviour):
  Model.compile / summary / fit / predict          scripts/train.py:227-235,308-316; scripts/inference.py:116
  load_model(path, custom_objects=, compile=False)  scripts/inference.py:218-226; scripts/benchmark.py:196-203
  optimizers.AdamW(learning_rate=, weight_decay=)   scripts/train.py:226
  metrics.MeanIoU(num_classes=, name=)              scripts/train.py:231; scripts/benchmark.py:237,269,277
  callbacks.{ModelCheckpoint, EarlyStopping, ReduceLROnPlateau, TensorBoard}   scripts/train.py:273-304
All arithmetic happens in libunet_b200.so kernels; this file is control flow, batching and host<->device copies.
"""
from __future__ import annotations

import math
import os
import sys
import time
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from .spec import UNetSpec

MODEL_NAME = "U-NET-Segmentation"      # model/u_net.py:114


class Scalar(float):
    """A host scalar that also answers `.numpy()`, like the eager tensors the reference's metric functions return."""

    def numpy(self):
        return np.float32(self)


# ====================================================================================================== optimizer
class AdamW:
    """tf.keras.optimizers.AdamW as used at scripts/train.py:226 (decoupled decay on every variable, eps 1e-7)."""

    def __init__(self, learning_rate: float = 1e-3, weight_decay: float = 4e-3, beta_1: float = 0.9,
                 beta_2: float = 0.999, epsilon: float = 1e-7, name: str = "AdamW"):
        self.learning_rate, self.weight_decay = float(learning_rate), float(weight_decay)
        self.beta_1, self.beta_2, self.epsilon, self.name = float(beta_1), float(beta_2), float(epsilon), name

    # ReduceLROnPlateau reads and writes `optimizer.learning_rate` / `.lr`
    @property
    def lr(self):
        return self.learning_rate

    @lr.setter
    def lr(self, v):
        self.learning_rate = float(v)

    def get_config(self):
        return dict(name=self.name, learning_rate=self.learning_rate, weight_decay=self.weight_decay,
                    beta_1=self.beta_1, beta_2=self.beta_2, epsilon=self.epsilon)


# ====================================================================================================== metrics
def _to_device_f32(a):
    import torch
    if isinstance(a, torch.Tensor):
        return a.to(device="cuda", dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float32))).cuda()


class MeanIoU:
    """tf.keras.metrics.MeanIoU(num_classes): confusion matrix over flattened labels / predictions cast to integers by
    truncation (so raw probabilities count as class 0 unless exactly 1.0 — the reference's train.py:231 behaviour),
    accumulated on the device in int64; result = mean over classes with a non-zero denominator of TP/(TP+FP+FN)."""

    def __init__(self, num_classes: int, name: str = "mean_io_u", dtype=None):
        if num_classes < 1 or num_classes > 64:
            raise ValueError("num_classes must be in 1..64")
        self.num_classes, self.name = int(num_classes), name
        self._counts = None

    def _dev_counts(self):
        import torch
        if self._counts is None:
            self._counts = torch.zeros(self.num_classes * self.num_classes, device="cuda", dtype=torch.int64)
        return self._counts

    def reset_state(self):
        if self._counts is not None:
            self._counts.zero_()

    reset_states = reset_state

    def update_state(self, y_true, y_pred, sample_weight=None):
        from . import ops
        if sample_weight is not None:
            raise NotImplementedError("sample_weight is not used by the reference and is not supported")
        t, p = _to_device_f32(y_true), _to_device_f32(y_pred)
        if t.numel() != p.numel():
            raise ValueError(f"y_true and y_pred differ in size: {tuple(t.shape)} vs {tuple(p.shape)}")
        ops.confusion_matrix_update(t.view(-1), p.view(-1), self.num_classes, self._dev_counts())

    def add_confusion(self, cm) -> None:
        """Accumulate a confusion matrix computed elsewhere (benchmark.py's device path sums per-sample counts)."""
        import torch
        c = torch.from_numpy(np.ascontiguousarray(np.asarray(cm, dtype=np.int64).reshape(-1))).cuda()
        if c.numel() != self.num_classes * self.num_classes:
            raise ValueError("confusion matrix has the wrong size")
        self._dev_counts().add_(c)

    def confusion_matrix(self) -> np.ndarray:
        c = self.num_classes
        return np.zeros((c, c), np.int64) if self._counts is None else self._counts.cpu().numpy().reshape(c, c)

    def result(self) -> Scalar:
        return Scalar(mean_iou_from_confusion(self.confusion_matrix()))


def mean_iou_from_confusion(cm: np.ndarray) -> float:
    cm = cm.astype(np.float64)
    tp = np.diag(cm)
    denom = cm.sum(0) + cm.sum(1) - tp
    valid = denom > 0
    if not valid.any():
        return 0.0
    return float((tp[valid] / denom[valid]).sum() / valid.sum())


def _loss_kind(loss) -> str:
    name = loss if isinstance(loss, str) else getattr(loss, "__name__", "")
    if name in ("dice_loss", "dice"):
        return "dice"
    if name in ("iou_loss", "jaccard_loss", "iou", "jaccard"):
        return "iou"
    raise ValueError(f"unsupported loss {loss!r}: the engine differentiates utils.loss.dice_loss and iou_loss/jaccard_loss")


# ====================================================================================================== callbacks
class Callback:
    def set_model(self, model):
        self.model = model

    def on_train_begin(self, logs=None): ...
    def on_train_end(self, logs=None): ...
    def on_epoch_begin(self, epoch, logs=None): ...
    def on_epoch_end(self, epoch, logs=None): ...


def _monitor_op(mode: str, monitor: str):
    if mode not in ("min", "max", "auto"):
        mode = "auto"
    if mode == "auto":
        mode = "max" if any(k in monitor for k in ("acc", "iou", "io_u", "dice_coef", "auc")) else "min"
    return (lambda a, b: a < b, math.inf) if mode == "min" else (lambda a, b: a > b, -math.inf)


class ModelCheckpoint(Callback):
    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", save_freq="epoch"):
        self.filepath, self.monitor, self.verbose = str(filepath), monitor, verbose
        self.save_best_only, self.save_weights_only = save_best_only, save_weights_only
        self.op, self.best = _monitor_op(mode, monitor)

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        path = self.filepath.format(epoch=epoch + 1, **logs)
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None:
                print(f"WARNING: Can save best model only with {self.monitor} available, skipping.")
                return
            if not self.op(cur, self.best):
                if self.verbose:
                    print(f"\nEpoch {epoch + 1}: {self.monitor} did not improve from {self.best:.5f}")
                return
            if self.verbose:
                print(f"\nEpoch {epoch + 1}: {self.monitor} improved from {self.best:.5f} to {cur:.5f}, saving model to {path}")
            self.best = cur
        elif self.verbose:
            print(f"\nEpoch {epoch + 1}: saving model to {path}")
        (self.model.save_weights if self.save_weights_only else self.model.save)(path)


class EarlyStopping(Callback):
    def __init__(self, monitor="val_loss", min_delta=0, patience=0, verbose=0, mode="auto", baseline=None,
                 restore_best_weights=False, start_from_epoch=0):
        self.monitor, self.patience, self.verbose = monitor, patience, verbose
        self.min_delta = abs(min_delta)
        self.restore_best_weights, self.start_from_epoch = restore_best_weights, start_from_epoch
        self.op, self.best = _monitor_op(mode, monitor)
        self._min = self.best == math.inf
        self.wait, self.stopped_epoch, self.best_epoch, self.best_weights = 0, 0, 0, None

    def on_train_begin(self, logs=None):
        self.wait, self.stopped_epoch, self.best_weights = 0, 0, None

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None or epoch < self.start_from_epoch:
            return
        if self.restore_best_weights and self.best_weights is None:
            self.best_weights = self.model.get_weights()
            self.best_epoch = epoch
        self.wait += 1
        improved = cur < self.best - self.min_delta if self._min else cur > self.best + self.min_delta
        if improved:
            self.best, self.best_epoch, self.wait = cur, epoch, 0
            if self.restore_best_weights:
                self.best_weights = self.model.get_weights()
            return
        if self.wait >= self.patience and epoch > 0:
            self.stopped_epoch = epoch
            self.model.stop_training = True

    def on_train_end(self, logs=None):
        if self.stopped_epoch > 0 and self.verbose:
            print(f"Epoch {self.stopped_epoch + 1}: early stopping")
        if self.restore_best_weights and self.best_weights is not None and self.stopped_epoch > 0:
            if self.verbose:
                print(f"Restoring model weights from the end of the best epoch: {self.best_epoch + 1}.")
            self.model.set_weights(self.best_weights)


class ReduceLROnPlateau(Callback):
    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", min_delta=1e-4,
                 cooldown=0, min_lr=0.0):
        if factor >= 1.0:
            raise ValueError("ReduceLROnPlateau does not support a factor >= 1.0.")
        self.monitor, self.factor, self.patience, self.verbose = monitor, factor, patience, verbose
        self.min_delta, self.cooldown, self.min_lr = min_delta, cooldown, min_lr
        _, self.best = _monitor_op(mode, monitor)
        self._min = self.best == math.inf
        self.wait, self.cooldown_counter = 0, 0

    def on_epoch_end(self, epoch, logs=None):
        logs = logs if logs is not None else {}
        opt = self.model.optimizer
        logs["learning_rate"] = opt.learning_rate
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        better = cur < self.best - self.min_delta if self._min else cur > self.best + self.min_delta
        if better:
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = opt.learning_rate
                if old > self.min_lr:
                    new = max(old * self.factor, self.min_lr)
                    opt.learning_rate = new
                    self.model._push_hyper()
                    if self.verbose:
                        print(f"\nEpoch {epoch + 1}: ReduceLROnPlateau reducing learning rate to {new}.")
                    self.cooldown_counter, self.wait = self.cooldown, 0


class TensorBoard(Callback):
    """Scalar summaries per epoch in TensorBoard event-file format (`tensorboard` package if importable, else a CSV in
    the same directory).  histogram_freq = n > 0 (scripts/train.py:302 passes 1): every n-th epoch the weights of every layer
    are written as histogram summaries under the Keras tag `<layer>/<weight>` (train run)."""

    def __init__(self, log_dir="logs", histogram_freq=0, **_):
        self.log_dir, self.histogram_freq = str(log_dir), histogram_freq
        self._writers = {}

    def _writer(self, sub):
        if sub not in self._writers:
            path = os.path.join(self.log_dir, sub)
            os.makedirs(path, exist_ok=True)
            try:
                from tensorboard.summary.writer.event_file_writer import EventFileWriter
                self._writers[sub] = ("tb", EventFileWriter(path))
            except Exception:
                self._writers[sub] = ("csv", open(os.path.join(path, "scalars.csv"), "a"))
        return self._writers[sub]

    def on_epoch_end(self, epoch, logs=None):
        for k, v in (logs or {}).items():
            sub, tag = ("validation", "epoch_" + k[4:]) if k.startswith("val_") else ("train", "epoch_" + k)
            kind, w = self._writer(sub)
            if kind == "tb":
                from tensorboard.compat.proto.event_pb2 import Event
                from tensorboard.compat.proto.summary_pb2 import Summary
                w.add_event(Event(wall_time=time.time(), step=epoch,
                                  summary=Summary(value=[Summary.Value(tag=tag, simple_value=float(v))])))
            else:
                w.write(f"{epoch},{tag},{float(v)}\n")
        if self.histogram_freq and (epoch + 1) % int(self.histogram_freq) == 0:
            self._write_histograms(epoch)

    def _write_histograms(self, epoch):
        kind, w = self._writer("train")
        if kind != "tb":
            return
        from tensorboard.compat.proto.event_pb2 import Event
        from tensorboard.compat.proto.summary_pb2 import HistogramProto, Summary
        values = []
        for name, arr in self.model.get_weights_dict().items():
            a = np.asarray(arr, np.float64).reshape(-1)
            counts, edges = np.histogram(a, bins=30)
            h = HistogramProto(min=float(a.min()), max=float(a.max()), num=int(a.size), sum=float(a.sum()),
                               sum_squares=float((a * a).sum()), bucket_limit=[float(e) for e in edges[1:]],
                               bucket=[float(c) for c in counts])
            values.append(Summary.Value(tag=name, histo=h))
        w.add_event(Event(wall_time=time.time(), step=epoch, summary=Summary(value=values)))

    def on_train_end(self, logs=None):
        for kind, w in self._writers.values():
            (w.close if kind == "tb" else w.close)()
        self._writers = {}


class History(Callback):
    def __init__(self):
        self.history: Dict[str, List[float]] = {}
        self.epoch: List[int] = []

    def on_epoch_end(self, epoch, logs=None):
        self.epoch.append(epoch)
        for k, v in (logs or {}).items():
            self.history.setdefault(k, []).append(v)


# ====================================================================================================== layers (names are API)
class Layer:
    def __init__(self, model, info):
        self._model, self.name, self.kind = model, info.name, info.kind
        self.output_shape, self._params, self.connected_to = info.out_shape, info.params, info.connected_to

    def count_params(self) -> int:
        return self._params

    @property
    def weight_names(self) -> List[str]:
        return self._model.spec.layer_weight_names(self.name)

    def get_weights(self) -> List[np.ndarray]:
        w = self._model.get_weights_dict()
        return [w[n] for n in self.weight_names]

    def set_weights(self, arrays: Sequence[np.ndarray]) -> None:
        names = self.weight_names
        if len(arrays) != len(names):
            raise ValueError(f"layer {self.name} expects {len(names)} weight arrays, got {len(arrays)}")
        self._model.set_weights_dict(dict(zip(names, arrays)))


# ====================================================================================================== input prefetch
_COPY_THREADS = None


def _copy_threads() -> int:
    """torch's CPU copy kernel is the staging memcpy (measured on the B200 host: 12.8 GB/s on one thread, 50 GB/s on 16; a
    Python thread pool over np.copyto is slower than ONE thread).  torchrun exports OMP_NUM_THREADS=1, which would leave a 268 MB
    batch to a single thread: on first use give torch this process's share of the host cores."""
    global _COPY_THREADS
    if _COPY_THREADS is None:
        import torch
        try:
            cores = len(os.sched_getaffinity(0))
        except Exception:
            cores = os.cpu_count() or 1
        local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1))
        want = max(1, min(32, cores // local_world))
        if torch.get_num_threads() < want:
            torch.set_num_threads(want)
        _COPY_THREADS = torch.get_num_threads()
    return _COPY_THREADS


def _host_copy(dst, src) -> None:
    """Host-to-host copy (with dtype conversion) between NumPy arrays / CPU tensors on all of this process's CPU threads: one
    thread moves ~10 GB/s, which would cap `predict` of 512x512 batches below what the GPU and PCIe sustain."""
    import torch
    nbytes = src.numel() * src.element_size() if isinstance(src, torch.Tensor) else src.nbytes
    if nbytes < (8 << 20):                             # small batches: the thread fan-out costs more than it saves
        np.copyto(dst.numpy() if isinstance(dst, torch.Tensor) else dst, src.numpy() if isinstance(src, torch.Tensor) else src,
                  casting="unsafe")
        return
    _copy_threads()
    try:
        d = dst if isinstance(dst, torch.Tensor) else torch.from_numpy(dst)
        s_ = src if isinstance(src, torch.Tensor) else torch.from_numpy(src if src.flags.writeable else src.copy())
        d.copy_(s_)
    except (TypeError, ValueError, RuntimeError):      # dtypes / layouts torch cannot wrap: NumPy does it on one thread
        np.copyto(dst.numpy() if isinstance(dst, torch.Tensor) else dst, src.numpy() if isinstance(src, torch.Tensor) else src,
                  casting="unsafe")


class _Prefetcher:
    """Host -> device staging for `fit`: batch i+1 is uploaded on a copy stream (from pinned memory) while batch i
    computes.  Two device slots; a slot is rewritten only after the step that read it has been enqueued and has
    finished on the compute stream (event), and a step starts only after its upload has landed (event)."""

    def __init__(self, source, cache: Optional[dict] = None, with_y: bool = True):
        import torch
        self.src = iter(source)
        self.with_y = with_y
        cache = cache if cache is not None else {}
        if "stream" not in cache:           # the copy stream and the staging buffers outlive one fit()/evaluate() call
            cache["stream"] = torch.cuda.Stream()
            cache["slots"] = [dict(dev={}, pin={}, ready=None, done=None) for _ in range(2)]
        self.copy_stream, self.slots = cache["stream"], cache["slots"]
        self.h2d_bytes = 0

    def _upload(self, batch, slot):
        import torch
        x, y = batch[0], batch[1]
        out = []
        with torch.cuda.stream(self.copy_stream):
            if slot["done"] is not None:
                self.copy_stream.wait_event(slot["done"])
            for key, arr in ((("x", x), ("y", y)) if self.with_y else (("x", x),)):
                if isinstance(arr, torch.Tensor) and arr.is_cuda:
                    out.append(arr.to(torch.float32).contiguous())
                    continue
                if isinstance(arr, torch.Tensor) and arr.is_pinned() and arr.dtype == torch.float32 and arr.is_contiguous():
                    host = arr
                else:
                    a = arr.numpy() if isinstance(arr, torch.Tensor) else np.asarray(arr)
                    host = slot["pin"].get((key, a.shape))
                    if host is None:
                        host = slot["pin"][(key, a.shape)] = torch.empty(a.shape, dtype=torch.float32).pin_memory()
                    if slot["ready"] is not None:
                        slot["ready"].synchronize()          # the previous DMA out of this pinned buffer has finished
                    _host_copy(host, a)
                dev = slot["dev"].get((key, tuple(host.shape)))
                if dev is None:
                    dev = slot["dev"][(key, tuple(host.shape))] = torch.empty(host.shape, dtype=torch.float32, device="cuda")
                dev.copy_(host, non_blocking=True)
                self.h2d_bytes += host.numel() * 4
                out.append(dev)
            slot["ready"] = torch.cuda.Event()
            slot["ready"].record(self.copy_stream)
        return out

    def __iter__(self):
        import torch
        i = 0
        try:
            nxt = self._upload(next(self.src), self.slots[0])
        except StopIteration:
            return
        while nxt is not None:
            cur, cur_slot = nxt, self.slots[i % 2]
            try:
                nxt = self._upload(next(self.src), self.slots[(i + 1) % 2])
            except StopIteration:
                nxt = None
            torch.cuda.current_stream().wait_event(cur_slot["ready"])
            yield cur[0], (cur[1] if self.with_y else None)
            cur_slot["done"] = torch.cuda.Event()
            cur_slot["done"].record(torch.cuda.current_stream())
            i += 1


# ====================================================================================================== model
class Model:
    """What `U_NET(...)` returns.  The engine (device buffers, kernels) is created on first use, so constructing the
    model, `summary()` and weight bookkeeping work without a GPU; running it does not."""

    def __init__(self, input_size, num_classes=1, dropout_rate=0.2, use_batch_norm=True, dtype: Optional[str] = None,
                 seed: int = 2301):
        self.spec = UNetSpec(tuple(input_size), int(num_classes), float(dropout_rate), bool(use_batch_norm))
        self.name = MODEL_NAME
        self.dtype_name = dtype or os.environ.get("UNET_B200_DTYPE", "bf16")
        self.layers = [Layer(self, li) for li in self.spec.layers]
        self.input_shape = (None,) + tuple(self.spec.input_size)
        self.output_shape = (None,) + tuple(self.spec.input_size[:2]) + (self.spec.num_classes,)
        self.optimizer: Optional[AdamW] = None
        self.loss = None
        self.metrics: list = []
        self.stop_training = False
        self._engine = None
        self._seed = seed
        self._pending_weights: Dict[str, np.ndarray] = {}
        self._grad_sync = None
        self._pinned = {}

    # ---------------------------------------------------------------- engine / weights
    @property
    def engine(self):
        if self._engine is None:
            from .engine import UNetEngine
            sp = self.spec
            self._engine = UNetEngine(sp.input_size, sp.num_classes, sp.dropout_rate, sp.use_batch_norm,
                                      dtype=self.dtype_name, seed=self._seed)
            if self._pending_weights:
                self._engine.set_weights(self._pending_weights)
                self._pending_weights = {}
            self._engine.use_graphs = os.environ.get("UNET_B200_GRAPHS", "1") != "0"
            self._push_hyper()
        return self._engine

    def get_layer(self, name: str) -> Layer:
        for l in self.layers:
            if l.name == name:
                return l
        raise ValueError(f"No such layer: {name}. Existing layers are: {[l.name for l in self.layers]}.")

    def count_params(self) -> int:
        return self.spec.trainable_params + self.spec.non_trainable_params

    def get_weights_dict(self) -> Dict[str, np.ndarray]:
        return self.engine.get_weights()

    def set_weights_dict(self, weights: Dict[str, np.ndarray]) -> None:
        for name, a in weights.items():
            if name not in self.spec.params:
                raise KeyError(f"unknown weight {name!r}")
            if tuple(np.shape(a)) != tuple(self.spec.params[name].shape):
                raise ValueError(f"{name}: expected shape {self.spec.params[name].shape}, got {np.shape(a)}")
        if self._engine is None:
            self._pending_weights.update({k: np.asarray(v, np.float32) for k, v in weights.items()})
        else:
            self._engine.set_weights(weights)

    def get_weights(self) -> List[np.ndarray]:
        d = self.get_weights_dict()
        return [d[n] for n in self.spec.params]          # Keras order: per layer, in creation order

    def set_weights(self, arrays: Sequence[np.ndarray]) -> None:
        names = list(self.spec.params)
        if len(arrays) != len(names):
            raise ValueError(f"expected {len(names)} weight arrays, got {len(arrays)}")
        self.set_weights_dict(dict(zip(names, arrays)))

    # ---------------------------------------------------------------- compile / summary
    def compile(self, optimizer=None, loss=None, metrics=None, **_):
        if isinstance(optimizer, str):
            if optimizer.lower() != "adamw":
                raise ValueError("only the AdamW optimizer of the reference (train.py:226) is implemented")
            optimizer = AdamW()
        self.optimizer = optimizer
        self.loss = loss
        self._loss_kind = _loss_kind(loss) if loss is not None else None
        self.metrics = list(metrics or [])
        for m in self.metrics:
            if not isinstance(m, MeanIoU) and getattr(m, "__name__", m) not in ("dice_coef", "iou_coef"):
                raise ValueError(f"unsupported metric {m!r}")
        if self._engine is not None:
            self._push_hyper()

    def _push_hyper(self):
        if self._engine is not None and self.optimizer is not None:
            o = self.optimizer
            self._engine.set_hyper(lr=o.learning_rate, weight_decay=o.weight_decay, beta1=o.beta_1, beta2=o.beta_2,
                                   eps=o.epsilon)

    def summary(self, line_length: int = 100, print_fn: Callable[[str], None] = print):
        ll = max(int(line_length or 100), 60)
        cols = [int(ll * 0.40), int(ll * 0.30), int(ll * 0.12)]

        def row(a, b, c, d):
            s = a[:cols[0] - 1].ljust(cols[0]) + b[:cols[1] - 1].ljust(cols[1]) + c[:cols[2] - 1].ljust(cols[2]) + d
            return s[:ll]

        print_fn(f'Model: "{self.name}"')
        print_fn("_" * ll)
        print_fn(row("Layer (type)", "Output Shape", "Param #", "Connected to"))
        print_fn("=" * ll)
        for l in self.layers:
            print_fn(row(f"{l.name} ({l.kind})", str(l.output_shape), f"{l.count_params():,}", l.connected_to))
        print_fn("=" * ll)
        tr, nt = self.spec.trainable_params, self.spec.non_trainable_params
        mb = lambda n: f"{n * 4 / 2 ** 20:.2f} MB"
        print_fn(f"Total params: {tr + nt:,} ({mb(tr + nt)})")
        print_fn(f"Trainable params: {tr:,} ({mb(tr)})")
        print_fn(f"Non-trainable params: {nt:,} ({mb(nt)})")
        print_fn("_" * ll)

    # ---------------------------------------------------------------- host <-> device staging
    def _stage_in(self, arr, key: str):
        """pinned host copy -> async H2D on the current stream; returns an fp32 device tensor."""
        import torch
        if isinstance(arr, torch.Tensor) and arr.is_cuda:
            return arr.to(torch.float32).contiguous()
        if isinstance(arr, torch.Tensor) and arr.is_pinned() and arr.dtype == torch.float32 and arr.is_contiguous():
            dev = self._pinned.get((key, "dev", tuple(arr.shape)))
            if dev is None:
                dev = self._pinned[(key, "dev", tuple(arr.shape))] = torch.empty(arr.shape, dtype=torch.float32, device="cuda")
            dev.copy_(arr, non_blocking=True)      # caller-owned pinned memory: straight DMA, no staging copy
            return dev
        a = arr.numpy() if isinstance(arr, torch.Tensor) else np.asarray(arr)
        slot = self._pinned.get((key, a.shape))
        if slot is None:
            slot = self._pinned[(key, a.shape)] = [torch.empty(a.shape, dtype=torch.float32).pin_memory(),
                                                   torch.empty(a.shape, dtype=torch.float32, device="cuda"), None]
        host, dev, ev = slot
        if ev is not None:
            ev.synchronize()          # the previous H2D out of this pinned buffer has landed (does not wait for compute)
        _host_copy(host, a)
        dev.copy_(host, non_blocking=True)
        slot[2] = torch.cuda.Event()
        slot[2].record()
        return dev

    # ---------------------------------------------------------------- inference
    def predict(self, x, batch_size: Optional[int] = 32, verbose=0, steps=None) -> np.ndarray:
        """model.predict (inference.py:116, benchmark.py:254): NHWC float array in, NHWC fp32 probabilities out.
        Chunks of `batch_size` are pipelined: upload of chunk i+1 (copy stream), forward of chunk i (compute stream) and
        download of chunk i-1 (second copy stream, into pinned memory) overlap; a worker thread copies finished chunks into
        the result array while the calling thread stages the next input."""
        import torch
        x = np.asarray(x) if not isinstance(x, torch.Tensor) else x
        if x.ndim != 4 or tuple(x.shape[1:]) != tuple(self.spec.input_size):
            raise ValueError(f'Input 0 of layer "{self.name}" is incompatible with the layer: expected shape='
                             f'(None, {", ".join(map(str, self.spec.input_size))}), found shape={tuple(x.shape)}')
        bs = int(batch_size or 32)
        n = x.shape[0]
        H, W = self.spec.input_size[:2]
        NC = self.spec.num_classes
        out = np.empty((n, H, W, NC), np.float32)
        if n == 0:
            return out
        eng = self.engine
        cache = self._pinned.setdefault(("predict", bs), {})
        if "in" not in cache:
            cache["in"] = torch.cuda.Stream(); cache["out"] = torch.cuda.Stream()
            cache["dev_out"] = [torch.empty((bs, H, W, NC), dtype=torch.float32, device="cuda") for _ in range(2)]
            cache["pin_out"] = [torch.empty((bs, H, W, NC), dtype=torch.float32).pin_memory() for _ in range(2)]
            cache["pf"] = {}
        if "worker" not in cache:
            from concurrent.futures import ThreadPoolExecutor
            cache["worker"] = ThreadPoolExecutor(1, thread_name_prefix="unet_predict_out")
        s_out = cache["out"]
        compute = torch.cuda.current_stream()
        dev_idx = torch.cuda.current_device()
        pending = [None, None]                       # per output slot: future of its copy-out job

        def copy_out(ev, slot, lo, hi):              # worker thread: the result array's first-touch page faults and the
            with torch.cuda.device(dev_idx):         # copy out of pinned memory stay off the thread that feeds the GPU
                ev.synchronize()
            _host_copy(out[lo:hi], cache["pin_out"][slot][: hi - lo])

        def retire(slot):
            if pending[slot] is not None:
                pending[slot].result()
                pending[slot] = None

        chunks = ((x[lo:lo + bs], None) for lo in range(0, n, bs))
        lo = 0
        for i, (xd, _) in enumerate(_Prefetcher(chunks, cache["pf"], with_y=False)):
            k = xd.shape[0]
            slot = i % 2
            probs = eng.forward_inference(xd)
            retire(slot)                             # the slot's previous download has been consumed by the host
            dev_out = cache["dev_out"][slot]
            dev_out[:k].copy_(probs[:k])             # device-to-device: frees the engine's output buffer for the next chunk
            ev_c = torch.cuda.Event(); ev_c.record(compute)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c)
                cache["pin_out"][slot][:k].copy_(dev_out[:k], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(s_out)
            pending[slot] = cache["worker"].submit(copy_out, ev, slot, lo, lo + k)
            lo += k
        retire(0); retire(1)
        return out

    __call__ = predict

    def predict_on_device(self, x):
        """Device tensor in, device probabilities out (plan-owned buffer, valid until the next call); no host copies."""
        return self.engine.forward_inference(x)

    # ---------------------------------------------------------------- training
    def _metric_names(self) -> List[str]:
        return [m.name if isinstance(m, MeanIoU) else m.__name__ for m in self.metrics]

    def enable_data_parallel(self, group=None):
        """Sum gradients across ranks (dist.GradSync) and scale by 1/world; average BN moving statistics."""
        from . import dist as D
        eng = self.engine
        self._grad_sync = D.GradSync(eng.g, D.grad_regions(eng.spec), group, mean_with_first=eng.state)
        eng.grad_hook = self._grad_sync.ready
        eng.grad_finish = self._grad_sync.finish
        eng.graph_collectives = os.environ.get("UNET_B200_GRAPH_NCCL", "1") != "0"
        eng.set_hyper(grad_scale=1.0 / self._grad_sync.world)

    def train_on_batch(self, x, y, return_dict=False):
        """One optimizer step; returns [loss, *metrics] as host floats (one device->host read)."""
        out3, extra = self._train_step_device(x, y)
        vals = out3.cpu().numpy()
        res = {"loss": float(vals[0])}
        for m in self.metrics:
            if isinstance(m, MeanIoU):
                res[m.name] = float(m.result())
            else:
                res[m.__name__] = float(vals[1] if m.__name__ == "dice_coef" else vals[2])
        return res if return_dict else list(res.values())

    def _train_step_device(self, x, y):
        if self.optimizer is None or self.loss is None:
            raise RuntimeError("You must call `compile()` before using the model for training.")
        eng = self.engine
        xd, yd = self._stage_in(x, "tx"), self._stage_in(y, "ty")
        out3 = eng.train_step(xd, yd, self._loss_kind)     # under data parallel the gradient / BN-statistics exchange is inside
        for m in self.metrics:
            if isinstance(m, MeanIoU):
                m.update_state(yd, eng._plans[(xd.shape[0], True)].t["probs"])
        return out3, None

    def _loss_ring(self):
        import torch
        if getattr(self, "_ring", None) is None:
            self._ring = torch.empty((8, 3), dtype=torch.float32).pin_memory()
        return self._ring

    def test_on_batch(self, x, y):
        eng = self.engine
        xd, yd = self._stage_in(x, "vx"), self._stage_in(y, "vy")
        out3 = eng.evaluate_batch(xd, yd, self._loss_kind or "dice")
        probs = eng._plans[(xd.shape[0], False)].t["probs"]
        return out3, yd, probs

    def _dp_reduce(self, acc, n, metrics):
        """Under data parallel every rank sees only its shard of each batch.  Sum the (loss, dice, iou) accumulators, the
        batch count and the MeanIoU confusion counts over the ranks, so that logs — and with them ModelCheckpoint,
        EarlyStopping and ReduceLROnPlateau decisions — are identical on every rank (ranks that disagreed about the learning
        rate would silently diverge; ranks that disagreed about stopping would hang in the next all-reduce)."""
        if self._grad_sync is None or self._grad_sync.world == 1:
            return acc, n
        import torch
        import torch.distributed as dist
        grp = self._grad_sync.group
        t = torch.cat([acc.double().to("cuda"), torch.tensor([float(n)], device="cuda", dtype=torch.float64)])
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=grp)
        for m in metrics:
            if isinstance(m, MeanIoU):
                dist.all_reduce(m._dev_counts(), op=dist.ReduceOp.SUM, group=grp)
        return t[:3], int(round(float(t[3].item())))

    def evaluate(self, x, y=None, steps=None, verbose=0, return_dict=False, _metrics=None):
        """Inference-mode loss / metrics over a generator of (x, y) batches (or arrays x, y)."""
        import torch
        batches = _iter_batches(x, y, steps)
        metrics = _metrics if _metrics is not None else self.metrics
        for m in metrics:
            if isinstance(m, MeanIoU):
                m.reset_state()
        acc = torch.zeros(3, device="cuda", dtype=torch.float64)
        n = 0
        for bx, by in batches:
            out3, yd, probs = self.test_on_batch(bx, by)
            acc += out3.double()
            n += 1
            for m in metrics:
                if isinstance(m, MeanIoU):
                    m.update_state(yd, probs)
        acc, n = self._dp_reduce(acc, n, metrics)      # data parallel: every rank reports the metrics of the WHOLE validation set
        vals = (acc / max(n, 1)).cpu().numpy()
        res = {"loss": float(vals[0])}
        for m in metrics:
            if isinstance(m, MeanIoU):
                res[m.name] = float(m.result())
            else:
                res[m.__name__] = float(vals[1] if m.__name__ == "dice_coef" else vals[2])
        return res if return_dict else list(res.values())

    def fit(self, x=None, y=None, batch_size=None, epochs=1, verbose=1, callbacks=None, validation_data=None,
            steps_per_epoch=None, validation_steps=None, initial_epoch=0, **_):
        """model.fit(generator, ...) of scripts/train.py:308-316.  One optimizer step per generator item; epoch logs
        carry loss, the compiled metrics and their val_ counterparts; callbacks see Keras' hooks."""
        import torch
        if self.optimizer is None or self.loss is None:
            raise RuntimeError("You must call `compile()` before using the model for training.")
        hist = History()
        cbs = [hist] + list(callbacks or [])
        for cb in cbs:
            cb.set_model(self)
        self.stop_training = False
        for cb in cbs:
            cb.on_train_begin()
        arrays = y is not None
        for epoch in range(initial_epoch, epochs):
            for cb in cbs:
                cb.on_epoch_begin(epoch)
            for m in self.metrics:
                if isinstance(m, MeanIoU):
                    m.reset_state()
            if arrays:
                it = _iter_batches(x, y, steps_per_epoch, batch_size or 32)
            else:
                if steps_per_epoch is None:
                    raise ValueError("steps_per_epoch is required when fitting from a generator")
                it = _take(x, steps_per_epoch)
            n = 0
            t0 = time.time()
            if verbose:
                print(f"Epoch {epoch + 1}/{epochs}")
            ring = self._loss_ring()
            host_sum = np.zeros(3, np.float64)
            pending = []                                   # (ring slot, event) of steps whose loss is still in flight
            pf = _Prefetcher(it, self._pinned.setdefault("prefetch", {}))
            for bx, by in pf:
                out3, _ = self._train_step_device(bx, by)
                slot = n % ring.shape[0]
                if len(pending) >= ring.shape[0] - 1:      # ring full: retire the oldest read first
                    k, ev = pending.pop(0); ev.synchronize(); host_sum += ring[k].numpy()
                ring[slot].copy_(out3, non_blocking=True)  # the step's device -> host read; never stalls the step
                ev = torch.cuda.Event(); ev.record()
                pending.append((slot, ev))
                n += 1
                while len(pending) > 1 and pending[0][1].query():
                    k, e0 = pending.pop(0); host_sum += ring[k].numpy()
                if verbose == 1 and (n % 10 == 0) and n > len(pending):
                    print(f"\r{n}/{steps_per_epoch or '?'} - loss: {host_sum[0] / (n - len(pending)):.4f}", end="", flush=True)
            for k, ev in pending:
                ev.synchronize(); host_sum += ring[k].numpy()
            self.last_h2d_bytes = pf.h2d_bytes
            self.engine.check_fold_guard()          # epoch end synchronises anyway (the logs are read back)
            if self._grad_sync is not None and self._grad_sync.world > 1:
                acc, n_all = self._dp_reduce(torch.tensor(host_sum, device="cuda", dtype=torch.float64), n, self.metrics)
                host_sum, n_red = acc.cpu().numpy(), n_all
            else:
                n_red = n
            vals = host_sum / max(n_red, 1)
            logs: Dict[str, float] = {"loss": float(vals[0])}
            for m in self.metrics:
                if isinstance(m, MeanIoU):
                    logs[m.name] = float(m.result())
                else:
                    logs[m.__name__] = float(vals[1] if m.__name__ == "dice_coef" else vals[2])
            if validation_data is not None:
                if isinstance(validation_data, tuple) and len(validation_data) == 2 and hasattr(validation_data[0], "shape"):
                    v = self.evaluate(validation_data[0], validation_data[1], return_dict=True)
                else:
                    if validation_steps is None:
                        raise ValueError("validation_steps is required when validating from a generator")
                    v = self.evaluate(validation_data, steps=validation_steps, return_dict=True)
                logs.update({"val_" + k: val for k, val in v.items()})
            if verbose:
                dt = time.time() - t0
                print(("\r" if verbose == 1 else "") + f"{n}/{n} - {dt:.0f}s - " +
                      " - ".join(f"{k}: {val:.4f}" for k, val in logs.items()))
            for cb in cbs:
                cb.on_epoch_end(epoch, logs)
            if self.stop_training:
                break
        for cb in cbs:
            cb.on_train_end()
        return hist

    # ---------------------------------------------------------------- serialization
    def get_config(self) -> dict:
        sp = self.spec
        return dict(name=self.name, input_size=list(sp.input_size), num_classes=sp.num_classes,
                    dropout_rate=sp.dropout_rate, use_batch_norm=sp.use_batch_norm)

    def save(self, filepath, overwrite=True, **_):
        from . import weights_io
        weights_io.save_model(self, str(filepath))

    def save_weights(self, filepath, overwrite=True, **_):
        from . import weights_io
        weights_io.save_model(self, str(filepath), weights_only=True)

    def load_weights(self, filepath, **_):
        from . import weights_io
        self.set_weights_dict(weights_io.read_weights(str(filepath), self.spec))


def _take(gen: Iterable, n: int):
    it = iter(gen)
    for _ in range(n):
        try:
            yield next(it)
        except StopIteration:
            return


def _iter_batches(x, y, steps, batch_size: int = 32):
    if y is None:
        if steps is None:
            raise ValueError("steps is required when evaluating from a generator")
        yield from _take(x, steps)
        return
    n = len(x)
    k = 0
    for lo in range(0, n, batch_size):
        if steps is not None and k >= steps:
            return
        yield x[lo:lo + batch_size], y[lo:lo + batch_size]
        k += 1


def load_model(filepath, custom_objects=None, compile=False, dtype: Optional[str] = None, **_) -> Model:
    """tf.keras.models.load_model(path, custom_objects=..., compile=False) (inference.py:226, benchmark.py:203):
    rebuilds the U-Net from the file's model configuration and loads its weights."""
    from . import weights_io
    return weights_io.load_model(str(filepath), dtype=dtype)
