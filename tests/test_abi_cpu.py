"""The drop-in boundary without a GPU: the shared library builds, loads, and exports exactly what include/unet_b200.h
declares; the ctypes table (unet_b200/_lib.py) agrees with the header on every name and argument count; the host-only
entry points work; compute entry points refuse to run rather than fall back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "unet_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"^(?:int|const char\*|uint32_t)\s+(unet_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.M | re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    return out


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from unet_b200 import _lib
    return _lib


def test_header_and_binding_agree(lib):
    hdr = header_functions()
    assert len(hdr) >= 27
    assert set(hdr) == set(lib.EXPORTS), set(hdr) ^ set(lib.EXPORTS)
    for name, nargs in hdr.items():
        assert len(lib._SIGNATURES[name]) == nargs, name


def test_library_exports_every_declared_symbol(lib):
    so = ctypes.CDLL(str(lib.LIB_PATH))
    for name in header_functions():
        assert hasattr(so, name), name
    assert lib.load().unet_sm_arch() == 100
    assert lib.load().unet_version() >= 100


def test_sass_is_blackwell_native(lib):
    """The built library carries sm_100a code only, with tcgen05 / TMA instructions in it."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elf = subprocess.run([cuobjdump, "-lelf", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", elf))
    assert archs == {"100a"}, archs
    sass = subprocess.run([cuobjdump, "-sass", str(lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCQMMA" in sass          # tcgen05.mma
    assert "LDTM" in sass                                  # tcgen05.ld
    assert "UTMALDG" in sass and "UTMASTG" in sass          # TMA loads and stores


def test_host_dropout_hash_matches_oracle(lib):
    from oracle import unet_ref as R
    l = lib.load()
    idx = np.array([0, 1, 2, 12345, 2 ** 32 + 5, 2 ** 40 + 77], dtype=np.uint64)
    for seed in (0, 7, 2 ** 31 + 3):
        got = [l.unet_host_dropout_hash(int(i), seed) for i in idx]
        assert got == R.dropout_hash(idx, seed).tolist()


def test_no_cpu_fallback(lib):
    """Without a GPU every compute path fails loudly: the engine refuses to construct and a raw kernel call errors."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from unet_b200.engine import UNetEngine
    with pytest.raises(RuntimeError):
        UNetEngine((32, 32, 3))
    rc = lib.load().unet_bn_fold(None, None, None, None, 1e-3, None, None, 8, None)
    assert rc != 0 and b"bn_fold" in lib.load().unet_last_error()
    import pathlib
    for mod in ["engine.py", "ops.py", "keras_api.py", "dist.py", "weights_io.py", "imaging.py", "data.py", "h5lite.py"]:
        text = (pathlib.Path(ROOT) / "unet-image-segmentation_b200" / mod).read_text()
        assert "import oracle" not in text and "from oracle" not in text, mod
    for mod in ["model/u_net.py", "utils/loss.py", "utils/metrics.py", "scripts/train.py", "scripts/inference.py", "scripts/benchmark.py"]:
        text = (pathlib.Path(ROOT) / mod).read_text()
        assert "import oracle" not in text and "from oracle" not in text, mod
