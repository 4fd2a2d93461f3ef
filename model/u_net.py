"""Drop-in for the reference `model/u_net.py`: same builder names, signatures, layer names, printed progress and
error behaviour (reference model/u_net.py:5-116), but the returned model executes on hand-written sm_100a kernels
(unet_b200.engine) instead of a Keras/TensorFlow graph.

    from model.u_net import U_NET
    model = U_NET((256, 256, 3), num_classes=1, dropout_rate=0.2, use_batch_norm=True)

`conv_block` exists for API completeness: in the reference it wires three Keras layers into a functional graph; here
the topology is data (unet_b200.spec), so it returns the layer descriptors a block would add.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

from unet_b200.keras_api import Model
from unet_b200.spec import FILTERS


def conv_block(input_tensor=None, num_filters: int = 64, kernel_size: int = 3, use_batch_norm: bool = True,
               name_prefix: Optional[str] = None) -> List[Tuple[str, str]]:
    """SeparableConv2D(num_filters, kernel_size, padding='same', use_bias=not use_batch_norm) -> [BatchNormalization]
    -> Activation('relu')  (reference model/u_net.py:5-26).  Returns [(layer name, Keras class)] of the block."""
    if kernel_size != 3:
        raise ValueError("the B200 engine implements the reference's 3x3 separable blocks only")
    p = name_prefix or "block"
    out = [(f"{p}_sepconv", "SeparableConv2D")]
    if use_batch_norm:
        out.append((f"{p}_bn", "BatchNormalization"))
    out.append((f"{p}_relu", "Activation"))
    return out


def U_NET(input_size: Tuple[int, int, int], num_classes: int = 1, dropout_rate: float = 0.2,
          use_batch_norm: bool = True) -> Model:
    """Builds the U-Net of the reference (4 encoder stages 64/128/256/512, 1024-filter bottleneck, 4 decoder stages with
    Conv2DTranspose + skip concatenation, 1x1 sigmoid/softmax head).  Raises ValueError unless `input_size` is
    (height, width, channels) (reference model/u_net.py:52-53)."""
    if len(input_size) != 3:
        raise ValueError("input_size must be a tuple of (height, width, channels)")
    filters = list(FILTERS)
    print("Building Encoder...")
    for i, f in enumerate(filters):
        print(f"  Encoder Stage {i + 1}, Filters: {f}")
    print("Building Bottleneck...")
    print(f"  Bottleneck Filters: {filters[-1] * 2}")
    print("Building Decoder...")
    for i, f in enumerate(reversed(filters)):
        print(f"  Decoder Stage {len(filters) - i}, Filters: {f}")
    print("Building Output Layer...")
    model = Model(tuple(input_size), num_classes=num_classes, dropout_rate=dropout_rate, use_batch_norm=use_batch_norm)
    print("U-Net model built successfully.")
    return model
