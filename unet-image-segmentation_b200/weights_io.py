"""Model files: what `ModelCheckpoint` writes (scripts/train.py:273-280) and `load_model` reads
(scripts/inference.py:226, scripts/benchmark.py:203).

Formats, chosen by file content when reading and by extension when writing:
  *.h5 / *.hdf5   Keras legacy full-model HDF5 (root attrs model_config / keras_version / backend, group
                  `model_weights/<layer>/<layer>/<weight>:0` with `layer_names` / `weight_names` attrs) through the
                  dependency-free HDF5 subset in h5lite.py
  *.keras         Keras-3 zip: config.json + metadata.json + model.weights.h5 (`layers/<name>/vars/<i>`)
  *.npz           name -> array, plus `__config__` (JSON)
"""
from __future__ import annotations

import io
import json
import os
import zipfile
from typing import Dict, Optional, Tuple

import numpy as np

from .spec import UNetSpec

KERAS_VERSION = "2.15.0"


def _model_config(sp: UNetSpec, name: str = "U-NET-Segmentation") -> dict:
    """A Keras-style functional-model config: enough for our own loader (the `unet_b200` block) and for a human
    reading the file; layer class names and names follow model/u_net.py."""
    return {
        "class_name": "Functional",
        "config": {
            "name": name,
            "layers": [{"class_name": l.kind, "name": l.name, "inbound_nodes": l.connected_to} for l in sp.layers],
            "input_layers": [["input_image", 0, 0]],
            "output_layers": [["output_mask", 0, 0]],
        },
        "unet_b200": {"input_size": list(sp.input_size), "num_classes": sp.num_classes,
                      "dropout_rate": sp.dropout_rate, "use_batch_norm": sp.use_batch_norm},
    }


def spec_from_config(cfg: dict) -> UNetSpec:
    """Recover U_NET(...) arguments from a model config: our own block if present, else from the Keras layer list
    (input shape from the InputLayer, classes from `output_mask`, BN from the presence of `*_bn` layers, dropout rate
    from `bneck_dropout`)."""
    if "unet_b200" in cfg:
        u = cfg["unet_b200"]
        return UNetSpec(tuple(u["input_size"]), u["num_classes"], u["dropout_rate"], u["use_batch_norm"])
    layers = cfg.get("config", {}).get("layers", [])
    by_name = {l.get("name") or l.get("config", {}).get("name"): l for l in layers}
    inp = by_name.get("input_image")
    if inp is None:
        raise ValueError("model file does not describe the reference U-Net (no `input_image` layer)")
    ic = inp.get("config", {})
    shape = ic.get("batch_input_shape") or ic.get("batch_shape")
    if not shape:
        raise ValueError("model config has no input shape")
    out = by_name.get("output_mask", {}).get("config", {})
    drop = by_name.get("bneck_dropout", {}).get("config", {}).get("rate", 0.0)
    return UNetSpec(tuple(shape[1:]), int(out.get("filters", 1)), float(drop), "enc1_block1_bn" in by_name)


# ------------------------------------------------------------------------------------------------ writers
def _weights_by_layer(spec: UNetSpec, w: Dict[str, np.ndarray]) -> Dict[str, Dict[str, np.ndarray]]:
    out: Dict[str, Dict[str, np.ndarray]] = {}
    for name in spec.params:
        layer, leaf = name.split("/")
        out.setdefault(layer, {})[leaf] = np.asarray(w[name], np.float32)
    return out


def save_model(model, path: str, weights_only: bool = False) -> None:
    write_model_file(path, model.spec, model.get_weights_dict(), model.name, weights_only)


def write_model_file(path: str, spec: UNetSpec, weights: Dict[str, np.ndarray], name: str = "U-NET-Segmentation",
                     weights_only: bool = False) -> None:
    ext = os.path.splitext(path)[1].lower()
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    cfg = _model_config(spec, name)
    if ext == ".npz":
        arrays = dict(weights)
        arrays["__config__"] = np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8)
        with open(path, "wb") as f:
            np.savez(f, **arrays)
        return
    from . import h5lite
    if ext == ".keras":
        root = h5lite.Group()
        layers = root.group("layers")
        for layer, ws in _weights_by_layer(spec, weights).items():
            vars_ = layers.group(layer).group("vars")
            for i, (leaf, arr) in enumerate(ws.items()):
                vars_.dataset(str(i), arr)
        buf = io.BytesIO()
        h5lite.write(buf, root)
        with zipfile.ZipFile(path, "w", zipfile.ZIP_STORED) as z:
            z.writestr("config.json", json.dumps(cfg))
            z.writestr("metadata.json", json.dumps({"keras_version": "3.0.0", "date_saved": ""}))
            z.writestr("model.weights.h5", buf.getvalue())
        return
    # legacy HDF5 (.h5, .hdf5, anything else — Keras also treats unknown suffixes as HDF5 in 2.x)
    root = h5lite.Group()
    mw = root if weights_only else root.group("model_weights")
    by_layer = _weights_by_layer(spec, weights)
    layer_names = [l.name for l in spec.layers]
    mw.attrs["layer_names"] = [n.encode() for n in layer_names]
    mw.attrs["backend"] = b"tensorflow"
    mw.attrs["keras_version"] = KERAS_VERSION.encode()
    for lname in layer_names:
        g = mw.group(lname)
        ws = by_layer.get(lname, {})
        g.attrs["weight_names"] = [f"{lname}/{leaf}:0".encode() for leaf in ws]
        if ws:
            inner = g.group(lname)
            for leaf, arr in ws.items():
                inner.dataset(f"{leaf}:0", arr)
    if not weights_only:
        root.attrs["model_config"] = json.dumps(cfg).encode()
        root.attrs["keras_version"] = KERAS_VERSION.encode()
        root.attrs["backend"] = b"tensorflow"
    with open(path, "wb") as f:
        h5lite.write(f, root)


# ------------------------------------------------------------------------------------------------ readers
def _read_any(path: str) -> Tuple[Optional[dict], Dict[str, np.ndarray]]:
    """-> (model config or None, {"<layer>/<leaf>": array})."""
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic[:4] == b"\x89HDF":
        from . import h5lite
        with open(path, "rb") as f:
            root = h5lite.read(f.read())
        return _from_h5_legacy(root)
    if magic[:2] == b"PK":
        with zipfile.ZipFile(path) as z:
            names = z.namelist()
            if "model.weights.h5" in names:          # Keras-3 .keras
                from . import h5lite
                cfg = json.loads(z.read("config.json"))
                root = h5lite.read(z.read("model.weights.h5"))
                return cfg, _from_h5_keras3(root, cfg)
        with np.load(path) as z:                     # .npz
            cfg = json.loads(bytes(z["__config__"]).decode()) if "__config__" in z.files else None
            return cfg, {k: z[k] for k in z.files if k != "__config__"}
    raise ValueError(f"{path}: not an HDF5, .keras or .npz model file")


def _from_h5_legacy(root) -> Tuple[Optional[dict], Dict[str, np.ndarray]]:
    cfg = None
    mc = root.attrs.get("model_config")
    if mc is not None:
        cfg = json.loads(mc.decode() if isinstance(mc, bytes) else mc)
    mw = root.groups.get("model_weights", root)
    out: Dict[str, np.ndarray] = {}
    layer_names = mw.attrs.get("layer_names")
    if layer_names is None:
        layer_names = [n.encode() for n in mw.groups]
    for ln in layer_names:
        ln = ln.decode() if isinstance(ln, bytes) else ln
        g = mw.groups.get(ln)
        if g is None:
            continue
        for wn in g.attrs.get("weight_names", []):
            wn = wn.decode() if isinstance(wn, bytes) else wn
            node = g
            for part in wn.split("/"):
                node = node.groups[part] if part in node.groups else node.datasets[part]
            leaf = wn.split("/")[-1].split(":")[0]
            out[f"{ln}/{leaf}"] = np.asarray(node)
    return cfg, out


_K3_ORDER = {"SeparableConv2D": ["depthwise_kernel", "pointwise_kernel", "bias"],
             "BatchNormalization": ["gamma", "beta", "moving_mean", "moving_variance"],
             "Conv2DTranspose": ["kernel", "bias"], "Conv2D": ["kernel", "bias"]}


def _from_h5_keras3(root, cfg) -> Dict[str, np.ndarray]:
    spec = spec_from_config(cfg)
    out: Dict[str, np.ndarray] = {}
    layers = root.groups.get("layers", root)
    for li in spec.layers:
        names = spec.layer_weight_names(li.name)
        if not names:
            continue
        g = layers.groups.get(li.name)
        if g is None or "vars" not in g.groups:
            raise ValueError(f"weights file has no variables for layer {li.name}")
        vars_ = g.groups["vars"].datasets
        for i, n in enumerate(names):
            out[n] = np.asarray(vars_[str(i)])
    return out


def read_weights(path: str, spec: UNetSpec) -> Dict[str, np.ndarray]:
    _, w = _read_any(path)
    missing = [n for n in spec.params if n not in w]
    if missing:
        raise ValueError(f"{path}: missing weights {missing[:4]}{'...' if len(missing) > 4 else ''}")
    return {n: w[n] for n in spec.params}


def load_model(path: str, dtype: Optional[str] = None):
    from .keras_api import Model
    if not os.path.exists(path):
        raise OSError(f"No file or directory found at {path}")
    cfg, w = _read_any(path)
    if cfg is None:
        raise ValueError(f"{path} holds weights only (no model configuration); build U_NET(...) and call load_weights")
    spec = spec_from_config(cfg)
    m = Model(spec.input_size, spec.num_classes, spec.dropout_rate, spec.use_batch_norm, dtype=dtype)
    m.set_weights_dict({n: w[n] for n in spec.params})
    return m
