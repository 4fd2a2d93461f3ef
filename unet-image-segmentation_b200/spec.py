"""Topology of the reference U-Net (model/u_net.py:28-116) as data: Keras layer list, parameter table, flat layout.

Layer and weight names are API (Keras weight files are looked up by them):
  enc{s}_block{1,2}_{sepconv,bn,relu}, enc{s}_pool, bneck_block{1,2}_*, bneck_dropout,
  dec{s}_{upsample,concat,dropout}, dec{s}_block{1,2}_*, output_mask, input_image.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

FILTERS = (64, 128, 256, 512)     # u_net.py:57
BN_EPS = 1e-3                     # keras.layers.BatchNormalization defaults
BN_MOMENTUM = 0.99


@dataclass
class ConvBlock:
    """conv_block() of u_net.py:5-26: SeparableConv2D(3x3, same) -> [BatchNormalization] -> ReLU."""
    prefix: str
    cin: int
    cout: int
    level: int          # 0 = full resolution, 4 = bottleneck


@dataclass
class LayerInfo:
    name: str
    kind: str           # Keras class name
    out_shape: Tuple[Optional[int], ...]
    params: int
    connected_to: str = ""


@dataclass
class ParamInfo:
    name: str           # "<layer>/<weight>"
    shape: Tuple[int, ...]
    trainable: bool
    offset: int = 0     # element offset in the flat trainable / non-trainable buffer
    size: int = 0


def _align(n: int, a: int = 8) -> int:
    return (n + a - 1) // a * a


@dataclass
class UNetSpec:
    input_size: Tuple[int, int, int]
    num_classes: int = 1
    dropout_rate: float = 0.2
    use_batch_norm: bool = True
    blocks: List[ConvBlock] = field(default_factory=list)
    params: "Dict[str, ParamInfo]" = field(default_factory=dict)
    layers: List[LayerInfo] = field(default_factory=list)
    n_trainable_flat: int = 0
    n_state_flat: int = 0

    def __post_init__(self):
        if len(self.input_size) != 3:
            raise ValueError("input_size must be a tuple of (height, width, channels)")   # u_net.py:52-53
        self.input_size = tuple(int(v) for v in self.input_size)
        self._build()

    # ------------------------------------------------------------------ construction
    def _add_param(self, name, shape, trainable):
        size = 1
        for s in shape:
            size *= s
        if trainable:
            p = ParamInfo(name, tuple(shape), True, self.n_trainable_flat, size)
            self.n_trainable_flat = _align(self.n_trainable_flat + size)
        else:
            p = ParamInfo(name, tuple(shape), False, self.n_state_flat, size)
            self.n_state_flat = _align(self.n_state_flat + size)
        self.params[name] = p
        return size

    def _add_block(self, prefix, cin, cout, level, hw, prev):
        h, w = hw
        self.blocks.append(ConvBlock(prefix, cin, cout, level))
        n = self._add_param(f"{prefix}_sepconv/depthwise_kernel", (3, 3, cin, 1), True)
        n += self._add_param(f"{prefix}_sepconv/pointwise_kernel", (1, 1, cin, cout), True)
        if not self.use_batch_norm:
            n += self._add_param(f"{prefix}_sepconv/bias", (cout,), True)
        self.layers.append(LayerInfo(f"{prefix}_sepconv", "SeparableConv2D", (None, h, w, cout), n, prev))
        last = f"{prefix}_sepconv"
        if self.use_batch_norm:
            n = self._add_param(f"{prefix}_bn/gamma", (cout,), True)
            n += self._add_param(f"{prefix}_bn/beta", (cout,), True)
            n += self._add_param(f"{prefix}_bn/moving_mean", (cout,), False)
            n += self._add_param(f"{prefix}_bn/moving_variance", (cout,), False)
            self.layers.append(LayerInfo(f"{prefix}_bn", "BatchNormalization", (None, h, w, cout), n, last))
            last = f"{prefix}_bn"
        self.layers.append(LayerInfo(f"{prefix}_relu", "Activation", (None, h, w, cout), 0, last))
        return f"{prefix}_relu"

    def _build(self):
        h, w, cin = self.input_size
        self.layers.append(LayerInfo("input_image", "InputLayer", (None, h, w, cin), 0))
        prev = "input_image"
        c = cin
        hw = (h, w)
        skip_names = []
        for i, f in enumerate(FILTERS):
            s = i + 1
            prev = self._add_block(f"enc{s}_block1", c, f, i, hw, prev)
            prev = self._add_block(f"enc{s}_block2", f, f, i, hw, prev)
            skip_names.append(prev)
            hw = (hw[0] // 2, hw[1] // 2)
            self.layers.append(LayerInfo(f"enc{s}_pool", "MaxPooling2D", (None, hw[0], hw[1], f), 0, prev))
            prev = f"enc{s}_pool"
            c = f
        bf = FILTERS[-1] * 2
        prev = self._add_block("bneck_block1", c, bf, 4, hw, prev)
        prev = self._add_block("bneck_block2", bf, bf, 4, hw, prev)
        if self.dropout_rate > 0.0:
            self.layers.append(LayerInfo("bneck_dropout", "Dropout", (None, hw[0], hw[1], bf), 0, prev))
            prev = "bneck_dropout"
        c = bf
        for i, f in enumerate(reversed(FILTERS)):
            s = len(FILTERS) - i
            hw = (hw[0] * 2, hw[1] * 2)
            n = self._add_param(f"dec{s}_upsample/kernel", (2, 2, f, c), True)
            n += self._add_param(f"dec{s}_upsample/bias", (f,), True)
            self.layers.append(LayerInfo(f"dec{s}_upsample", "Conv2DTranspose", (None, hw[0], hw[1], f), n, prev))
            self.layers.append(LayerInfo(f"dec{s}_concat", "Concatenate", (None, hw[0], hw[1], 2 * f), 0,
                                         f"dec{s}_upsample, {skip_names[s - 1]}"))
            prev = f"dec{s}_concat"
            if self.dropout_rate > 0.0 and i < len(FILTERS) - 1:
                self.layers.append(LayerInfo(f"dec{s}_dropout", "Dropout", (None, hw[0], hw[1], 2 * f), 0, prev))
                prev = f"dec{s}_dropout"
            prev = self._add_block(f"dec{s}_block1", 2 * f, f, s - 1, hw, prev)
            prev = self._add_block(f"dec{s}_block2", f, f, s - 1, hw, prev)
            c = f
        n = self._add_param("output_mask/kernel", (1, 1, c, self.num_classes), True)
        n += self._add_param("output_mask/bias", (self.num_classes,), True)
        self.layers.append(LayerInfo("output_mask", "Conv2D", (None, hw[0], hw[1], self.num_classes), n, prev))

    # ------------------------------------------------------------------ queries
    @property
    def trainable_params(self) -> int:
        return sum(p.size for p in self.params.values() if p.trainable)

    @property
    def non_trainable_params(self) -> int:
        return sum(p.size for p in self.params.values() if not p.trainable)

    def layer_weight_names(self, layer: str) -> List[str]:
        """Weight names of one layer in Keras order."""
        return [n for n in self.params if n.split("/")[0] == layer]

    def has_dropout(self, name: str) -> bool:
        return any(l.name == name for l in self.layers)
