"""B200-native U-Net segmentation engine: hand-written sm_100a CUDA (csrc/) behind the C-ABI of
include/unet_b200.h, driven from Python.  Drop-in surface for planck-epoch/unet-image-segmentation:
`model.u_net.U_NET`, `utils.loss`, `utils.metrics`, `scripts/{train,inference,benchmark}.py`.

There is no CPU or library fallback: importing `unet_b200._lib` without the built shared library raises.
"""
__version__ = "0.1.0"
