"""Per-kernel parity: every C-ABI entry point of libunet_b200.so against the NumPy oracle (oracle/unet_ref.py) on the
same seeded inputs.  fp32 kernels: tight tolerances (summation order only).  bf16 kernels: operands are rounded to
bf16 on the host first, so the only differences are fp32 accumulation order and the final bf16 rounding (2^-8 rel).
"""
import numpy as np
import pytest
import torch

from oracle import unet_ref as R

pytestmark = pytest.mark.gpu

ops = None


@pytest.fixture(scope="module", autouse=True)
def _load():
    global ops
    from unet_b200 import ops as _ops
    _ops.device_check(0)
    ops = _ops


def dev(a, dtype=torch.float32):
    return torch.tensor(np.asarray(a), device="cuda").to(dtype).contiguous()


def host(t):
    torch.cuda.synchronize()
    return t.detach().float().cpu().numpy().astype(np.float64)


def bf16_round(a):
    return torch.tensor(np.asarray(a, dtype=np.float32)).to(torch.bfloat16).float().numpy().astype(np.float64)


def tol(dtype):
    return dict(rtol=2e-5, atol=2e-5) if dtype == torch.float32 else dict(rtol=1.0 / 128, atol=1e-2)


RNG = np.random.default_rng(2301)
DTYPES = [torch.float32, torch.bfloat16]


# ------------------------------------------------------------------------------------------------ depthwise
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 5, 7, 8), (3, 33, 9, 24), (2, 8, 8, 3), (1, 1, 1, 16), (1, 64, 48, 128),
                                   (2, 100, 70, 72), (1, 131, 33, 200), (2, 7, 5, 1), (1, 21, 64, 64), (2, 9, 130, 72), (1, 40, 32, 16)])
@pytest.mark.parametrize("flip", [False, True])
def test_dwconv_fwd(dtype, shape, flip):
    x = RNG.standard_normal(shape).astype(np.float32)
    w = RNG.standard_normal((3, 3, shape[3])).astype(np.float32)
    xr = bf16_round(x) if dtype == torch.bfloat16 else x.astype(np.float64)
    ref = R.dwconv3x3(xr, w[::-1, ::-1].astype(np.float64) if flip else w.astype(np.float64))
    xd, wd = dev(x, dtype), dev(w.reshape(9, -1))
    y = torch.empty_like(xd)
    ops.dwconv3x3(xd, wd, y, flip=flip)
    np.testing.assert_allclose(host(y), ref, **tol(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
def test_dwconv_fwd_views_affine_dropout(dtype):
    n, h, w, c, ctot = 2, 12, 10, 32, 96
    buf = RNG.standard_normal((n, h, w, ctot)).astype(np.float32)
    wk = RNG.standard_normal((3, 3, c)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, c).astype(np.float32)
    sh = RNG.standard_normal(c).astype(np.float32)
    bd = dev(buf, dtype)
    out = torch.zeros((n, h, w, 64), device="cuda", dtype=dtype)
    xin = bd[..., 40:40 + c]
    br = bf16_round(buf) if dtype == torch.bfloat16 else buf.astype(np.float64)
    xa = np.maximum(br[..., 40:40 + c] * sc + sh, 0)
    ref = R.dwconv3x3(xa, wk.astype(np.float64))
    drop = ops.make_dropout(0.25, 77, ctot=64, c0=16)
    ops.dwconv3x3(xin, dev(wk.reshape(9, -1)), out[..., 16:16 + c], in_scale=dev(sc), in_shift=dev(sh), drop=drop)
    mult = R.dropout_multiplier((n, h, w, 64), 0.25, 77)[..., 16:16 + c]
    got = host(out)
    np.testing.assert_allclose(got[..., 16:16 + c], ref * mult, **tol(dtype))
    assert np.all(got[..., :16] == 0) and np.all(got[..., 48:] == 0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 10, 66, 64), (1, 9, 32, 72), (2, 6, 18, 32)])
def test_dwconv_fwd_dropout_strip(dtype, shape):
    """Dropout on the stored output of the TMA-strip kernel (one and two columns per thread), written into a channel view"""
    n, h, w, c = shape
    x = RNG.standard_normal(shape).astype(np.float32)
    wk = RNG.standard_normal((3, 3, c)).astype(np.float32)
    xr = bf16_round(x) if dtype == torch.bfloat16 else x.astype(np.float64)
    ref = R.dwconv3x3(xr, wk.astype(np.float64))
    ctot = c + 24
    out = torch.zeros((n, h, w, ctot), device="cuda", dtype=dtype)
    ops.dwconv3x3(dev(x, dtype), dev(wk.reshape(9, -1)), out[..., 8:8 + c], drop=ops.make_dropout(0.25, 77, ctot=ctot, c0=8))
    mult = R.dropout_multiplier((n, h, w, ctot), 0.25, 77)[..., 8:8 + c]
    got = host(out)
    np.testing.assert_allclose(got[..., 8:8 + c], ref * mult, **tol(dtype))
    assert np.all(got[..., :8] == 0) and np.all(got[..., 8 + c:] == 0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 5, 7, 8), (2, 40, 9, 3), (1, 70, 12, 128), (2, 8, 8, 1024), (2, 33, 20, 1), (1, 9, 9, 4)])
def test_dwconv_bwd_weight(dtype, shape):
    x = RNG.standard_normal(shape).astype(np.float32)
    dy = RNG.standard_normal(shape).astype(np.float32)
    xr, dyr = (bf16_round(x), bf16_round(dy)) if dtype == torch.bfloat16 else (x.astype(np.float64), dy.astype(np.float64))
    _, ref = R.dwconv3x3_bwd(xr, np.zeros((3, 3, shape[3])), dyr)
    dw = torch.zeros((9, shape[3]), device="cuda")
    ops.dwconv3x3_bwd_weight(dev(x, dtype), dev(dy, dtype), dw)
    np.testing.assert_allclose(host(dw).reshape(3, 3, -1), ref, rtol=1e-4, atol=1e-3 * np.sqrt(np.prod(shape[:3])))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 5, 7, 8), (1, 70, 12, 128), (2, 8, 8, 1024), (2, 33, 40, 16), (1, 67, 35, 32),
                                   (1, 21, 96, 64), (2, 9, 50, 72)])
@pytest.mark.parametrize("mode", ["plain", "mask", "drop"])
def test_dwconv_bwd_fused(dtype, shape, mode):
    """one pass over dy: dx (+ReLU mask from x>0, BN-backward sums, or dropout on dx) and dw; ragged strips, views"""
    n, h, w, c = shape
    x = RNG.standard_normal(shape).astype(np.float32)
    if mode == "mask":
        x = np.maximum(x, 0)                            # x is a post-ReLU activation
    dy = RNG.standard_normal(shape).astype(np.float32)
    wk = RNG.standard_normal((3, 3, c)).astype(np.float32)
    xr, dyr = (bf16_round(x), bf16_round(dy)) if dtype == torch.bfloat16 else (x.astype(np.float64), dy.astype(np.float64))
    dx_ref, dw_ref = R.dwconv3x3_bwd(xr, wk.astype(np.float64), dyr)
    ctot = c + 16
    xbuf = torch.zeros((n, h, w, ctot), device="cuda", dtype=dtype); xbuf[..., 8:8 + c] = dev(x, dtype)
    dxbuf = torch.zeros((n, h, w, ctot), device="cuda", dtype=dtype)
    dw = torch.zeros((9, c), device="cuda")
    sums = torch.zeros((2, c), device="cuda") if mode == "mask" else None
    drop = ops.make_dropout(0.25, 91, ctot=ctot, c0=8) if mode == "drop" else None
    assert ops.dwconv3x3_bwd_supported(xbuf[..., 8:8 + c], dev(dy, dtype), dxbuf[..., 8:8 + c])
    ops.dwconv3x3_bwd(xbuf[..., 8:8 + c], dev(dy, dtype), dev(wk.reshape(9, -1)), dxbuf[..., 8:8 + c], dw,
                      relu_mask=(mode == "mask"), bn_sums=sums, drop=drop)
    if mode == "mask":
        dx_ref = dx_ref * (xr > 0)
    if mode == "drop":
        dx_ref = dx_ref * R.dropout_multiplier((n, h, w, ctot), 0.25, 91)[..., 8:8 + c]
    got = host(dxbuf)
    np.testing.assert_allclose(got[..., 8:8 + c], dx_ref, **tol(dtype))
    assert np.all(got[..., :8] == 0) and np.all(got[..., 8 + c:] == 0)
    np.testing.assert_allclose(host(dw).reshape(3, 3, -1), dw_ref, rtol=1e-4, atol=1e-3 * np.sqrt(n * h * w))
    if mode == "mask":
        g = got[..., 8:8 + c].astype(np.float64)
        np.testing.assert_allclose(host(sums)[0], g.sum((0, 1, 2)), rtol=1e-4, atol=1e-3 * np.sqrt(n * h * w))
        np.testing.assert_allclose(host(sums)[1], (g * xr).sum((0, 1, 2)), rtol=1e-4, atol=1e-3 * np.sqrt(n * h * w))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape,drop", [((2, 16, 32, 128), False), ((1, 10, 22, 256), True), ((2, 8, 6, 128), True), ((1, 34, 48, 256), False)])
def test_dwconv_bwd_fused_up_redirect(dtype, shape, drop):
    """the upsampled half of a concat gradient leaves un-pixel-shuffled (the Conv2DTranspose gradient operand) with its bias
    gradient: == the plain kernel followed by unet_convt_bwd_gather; the skip half still lands in dx"""
    n, h, w, c = shape
    f = c // 2
    x = RNG.standard_normal(shape).astype(np.float32); dy = RNG.standard_normal(shape).astype(np.float32)
    wk = RNG.standard_normal((3, 3, c)).astype(np.float32)
    xd, dyd, wd = dev(x, dtype), dev(dy, dtype), dev(wk.reshape(9, -1))
    dp = ops.make_dropout(0.25, 91, ctot=c, c0=0) if drop else None
    dx_ref = torch.empty(shape, device="cuda", dtype=dtype); dw_ref = torch.zeros((9, c), device="cuda")
    ops.dwconv3x3_bwd(xd, dyd, wd, dx_ref, dw_ref, drop=dp)
    g_ref = torch.empty((n * h * w // 4, 4 * f), device="cuda", dtype=dtype); db_ref = torch.zeros(f, device="cuda")
    ops.convt_bwd_gather(dx_ref[..., :f], g_ref, db_ref)
    dx = torch.full(shape, 7.0, device="cuda", dtype=dtype); dw = torch.zeros((9, c), device="cuda")
    g = torch.full_like(g_ref, 3.0); db = torch.zeros(f, device="cuda")
    ops.dwconv3x3_bwd(xd, dyd, wd, dx, dw, drop=dp, up_out=g, up_colsum=db)
    assert torch.equal(g, g_ref)
    assert torch.equal(dx[..., f:], dx_ref[..., f:]) and bool((dx[..., :f] == 7.0).all())
    np.testing.assert_allclose(host(dw), host(dw_ref), rtol=1e-4, atol=1e-3 * np.sqrt(n * h * w))
    # the bias gradient sums the fp32 values before they are rounded for storage
    xr, dyr = (bf16_round(x), bf16_round(dy)) if dtype == torch.bfloat16 else (x.astype(np.float64), dy.astype(np.float64))
    dxo, _ = R.dwconv3x3_bwd(xr, wk.astype(np.float64), dyr)
    if drop:
        dxo = dxo * R.dropout_multiplier((n, h, w, c), 0.25, 91)
    np.testing.assert_allclose(host(db), dxo[..., :f].sum((0, 1, 2)), rtol=1e-4, atol=1e-3 * np.sqrt(n * h * w))


@pytest.mark.parametrize("dtype", DTYPES)
def test_dwconv_bwd_fused_partial_dropout(dtype):
    """the Dropout mask is applied to channels >= drop_c_from only (the other half is masked by the reader of dx)"""
    n, h, w, c = 2, 9, 30, 128
    x = RNG.standard_normal((n, h, w, c)).astype(np.float32); dy = RNG.standard_normal((n, h, w, c)).astype(np.float32)
    wk = RNG.standard_normal((3, 3, c)).astype(np.float32)
    xr, dyr = (bf16_round(x), bf16_round(dy)) if dtype == torch.bfloat16 else (x.astype(np.float64), dy.astype(np.float64))
    dx_ref, _ = R.dwconv3x3_bwd(xr, wk.astype(np.float64), dyr)
    mult = R.dropout_multiplier((n, h, w, c), 0.25, 91)
    mult[..., :64] = 1.0
    dx = torch.empty((n, h, w, c), device="cuda", dtype=dtype); dw = torch.zeros((9, c), device="cuda")
    ops.dwconv3x3_bwd(dev(x, dtype), dev(dy, dtype), dev(wk.reshape(9, -1)), dx, dw,
                      drop=ops.make_dropout(0.25, 91, ctot=c, c0=0), drop_c_from=64)
    np.testing.assert_allclose(host(dx), dx_ref * mult, **tol(dtype))


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 16, 64), (1, 37, 45, 72), (2, 70, 33, 128), (1, 9, 100, 8), (2, 12, 64, 128), (1, 21, 98, 72),
                                   (1, 35, 32, 64)])
def test_dwconv_fwd_affine_on_load(dtype, shape):
    """TMA-strip kernel with the producer's BN+ReLU applied on load: zero padding lives in the transformed space"""
    n, h, w, c = shape
    z = RNG.standard_normal(shape).astype(np.float32)
    wk = RNG.standard_normal((3, 3, c)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, c).astype(np.float32); sh = (RNG.standard_normal(c) * 0.5 + 0.3).astype(np.float32)
    zr = bf16_round(z) if dtype == torch.bfloat16 else z.astype(np.float64)
    ref = R.dwconv3x3(np.maximum(zr * sc + sh, 0), wk.astype(np.float64))
    y = torch.empty(shape, device="cuda", dtype=dtype); cs = torch.zeros(c, device="cuda")
    ops.dwconv3x3(dev(z, dtype), dev(wk.reshape(9, -1)), y, in_scale=dev(sc), in_shift=dev(sh), colsum=cs)
    np.testing.assert_allclose(host(y), ref, **tol(dtype))
    np.testing.assert_allclose(host(cs), host(y).sum((0, 1, 2)), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mask", [False, True])
@pytest.mark.parametrize("shape", [(2, 37, 45, 72), (2, 37, 50, 72), (1, 19, 64, 128)])
def test_dwconv_bwd_fused_affine_on_load(dtype, mask, shape):
    """fused depthwise backward whose x stream is the producer's pre-BN tensor: y = relu(z*scale+shift) formed on load"""
    n, h, w, c = shape
    z = RNG.standard_normal((n, h, w, c)).astype(np.float32); dy = RNG.standard_normal((n, h, w, c)).astype(np.float32)
    wk = RNG.standard_normal((3, 3, c)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, c).astype(np.float32); sh = (RNG.standard_normal(c) * 0.5).astype(np.float32)
    zr, dyr = (bf16_round(z), bf16_round(dy)) if dtype == torch.bfloat16 else (z.astype(np.float64), dy.astype(np.float64))
    yact = np.maximum(zr * sc + sh, 0)
    dx_ref, dw_ref = R.dwconv3x3_bwd(yact, wk.astype(np.float64), dyr)
    if mask:
        dx_ref = dx_ref * (yact > 0)
    dx = torch.empty((n, h, w, c), device="cuda", dtype=dtype); dw = torch.zeros((9, c), device="cuda")
    sums = torch.zeros((2, c), device="cuda") if mask else None
    ops.dwconv3x3_bwd(dev(z, dtype), dev(dy, dtype), dev(wk.reshape(9, -1)), dx, dw, relu_mask=mask, bn_sums=sums,
                      x_scale=dev(sc), x_shift=dev(sh))
    np.testing.assert_allclose(host(dx), dx_ref, **tol(dtype))
    np.testing.assert_allclose(host(dw).reshape(3, 3, -1), dw_ref, rtol=1e-4, atol=1e-3 * np.sqrt(n * h * w))
    if mask:
        g = host(dx)
        np.testing.assert_allclose(host(sums)[0], g.sum((0, 1, 2)), rtol=1e-4, atol=1e-3 * np.sqrt(n * h * w))
        np.testing.assert_allclose(host(sums)[1], (g * yact).sum((0, 1, 2)), rtol=1e-4, atol=1e-3 * np.sqrt(n * h * w))


@pytest.mark.parametrize("shape", [(2, 37, 45, 72), (2, 19, 70, 72), (1, 33, 128, 64)])
def test_dwconv_fwd_colsum(shape):
    for dtype in DTYPES:
        x = RNG.standard_normal(shape).astype(np.float32)
        w = RNG.standard_normal((3, 3, shape[3])).astype(np.float32)
        y = torch.empty(shape, device="cuda", dtype=dtype)
        cs = torch.zeros(shape[3], device="cuda")
        ops.dwconv3x3(dev(x, dtype), dev(w.reshape(9, -1)), y, colsum=cs)
        np.testing.assert_allclose(host(cs), host(y).sum((0, 1, 2)), rtol=1e-4, atol=1e-2)


def test_gemm_tc_operand_concatenation():
    """[A | A2] along K (forward/dgrad layout) and [B | B2] along N (weight-gradient layout) == the concatenated operand"""
    bf = torch.bfloat16
    M, K1, K2, N = 1000, 64, 128, 96
    A, A2, Bt = dev(RNG.standard_normal((M, K1)), bf), dev(RNG.standard_normal((M, K2)), bf), dev(RNG.standard_normal((N, K1 + K2)), bf)
    bias = dev(RNG.standard_normal(N))
    ref = torch.empty((M, N), device="cuda", dtype=bf); got = torch.empty_like(ref)
    ops.gemm(torch.cat([A, A2], 1).contiguous(), Bt, ref, b_trans=True, epilogue=ops.EPI_AFFINE, shift=bias, tensor_core=True)
    ops.gemm(A, Bt, got, b_trans=True, A2=A2, epilogue=ops.EPI_AFFINE, shift=bias, tensor_core=True)
    assert torch.equal(ref, got)
    P, Mo, N1, N2 = 3000, 72, 128, 64
    Aw, B1, B2 = dev(RNG.standard_normal((P, Mo)), bf), dev(RNG.standard_normal((P, N1)), bf), dev(RNG.standard_normal((P, N2)), bf)
    ref = torch.zeros((Mo, N1 + N2), device="cuda"); got = torch.zeros_like(ref)
    ops.gemm(Aw, torch.cat([B1, B2], 1).contiguous(), ref, a_trans=True, accumulate=True, tensor_core=True)
    ops.gemm(Aw, B1, got, a_trans=True, accumulate=True, B2=B2, tensor_core=True)
    np.testing.assert_allclose(host(got), host(ref), rtol=1e-5, atol=1e-3)
    want = host(Aw).T @ np.concatenate([host(B1), host(B2)], 1)
    np.testing.assert_allclose(host(got), want, rtol=1e-3, atol=5e-2)


@pytest.mark.parametrize("cfg", [(1000, 64), (128 * 300 + 77, 128), (50, 64), (4096, 128)])
def test_pw_bwd_fused(cfg):
    """dd = [g | z] wab^T + bias and G += d^T [g | z] from one pass == the two tensor-core GEMMs (bit-exact dd) and NumPy"""
    bf = torch.bfloat16
    P, cin = cfg
    c = 64
    g, z = dev(RNG.standard_normal((P, c)), bf), dev(RNG.standard_normal((P, c)), bf)
    d = dev(RNG.standard_normal((P, cin)), bf)
    wab = dev(RNG.standard_normal((cin, 2 * c)) / np.sqrt(2 * c), bf)
    bias = dev(RNG.standard_normal(cin))
    dd_ref = torch.empty((P, cin), device="cuda", dtype=bf); G_ref = torch.zeros((cin, 2 * c), device="cuda")
    ops.gemm(g, wab, dd_ref, b_trans=True, A2=z, epilogue=ops.EPI_AFFINE, shift=bias, tensor_core=True)
    ops.gemm(d, g, G_ref, a_trans=True, accumulate=True, B2=z, tensor_core=True)
    assert ops.pw_bwd_fused_supported(g, z, d, dd_ref)
    # operands as column views of wider buffers (how the engine holds them), dd into a view as well
    gbuf = torch.zeros((P, c + 8), device="cuda", dtype=bf); gbuf[:, :c] = g
    ddbuf = torch.full((P, cin + 16), 7.0, device="cuda", dtype=bf)
    G = torch.zeros((cin, 2 * c), device="cuda")
    ops.pw_bwd_fused(gbuf[:, :c], z, d, wab, bias, ddbuf[:, 8:8 + cin], G)
    assert torch.equal(ddbuf[:, 8:8 + cin], dd_ref)
    assert bool((ddbuf[:, :8] == 7.0).all()) and bool((ddbuf[:, 8 + cin:] == 7.0).all())
    np.testing.assert_allclose(host(G), host(G_ref), rtol=1e-4, atol=1e-3 * np.sqrt(P))
    want = host(d).T @ np.concatenate([host(g), host(z)], 1)
    np.testing.assert_allclose(host(G), want, rtol=1e-3, atol=2e-3 * np.sqrt(P))
    ops.pw_bwd_fused(g, z, d, wab, bias, dd_ref, G)              # accumulates
    np.testing.assert_allclose(host(G), 2 * want, rtol=1e-3, atol=4e-3 * np.sqrt(P))


def test_bn_bwd_folded_into_gemms():
    """dz = A*g + B*z + K from the sums (sum g, sum g*y): folded data / weight gradients == explicit BatchNormalization backward"""
    bf = torch.bfloat16
    M, cin, c = 4096, 128, 64
    d = bf16_round(RNG.standard_normal((M, cin)))
    w = RNG.standard_normal((cin, c)).astype(np.float32) / np.sqrt(cin)
    z = bf16_round(d @ w.astype(np.float64) + 0.3)
    gamma = RNG.uniform(0.5, 1.5, c); beta = RNG.standard_normal(c) * 0.3
    mean, var = z.mean(0), z.var(0)
    rstd = 1.0 / np.sqrt(var + 1e-3)
    xhat = (z - mean) * rstd
    y = bf16_round(np.maximum(gamma * xhat + beta, 0))
    g = bf16_round(RNG.standard_normal((M, c)) * (y > 0))
    # explicit backward in fp64 (xhat from z, as unet_bn_bwd_reduce/apply do)
    dgamma, dbeta = (g * xhat).sum(0), g.sum(0)
    dz = gamma * rstd * (g - dbeta / M - xhat * dgamma / M)
    dd_ref, dw_ref = dz @ w.astype(np.float64).T, d.T @ dz
    sums = dev(np.stack([g.sum(0), (g * y).sum(0)]))
    dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    coef = torch.empty((3, c), device="cuda"); wab = torch.empty((cin, 2 * c), device="cuda", dtype=bf); bias = torch.empty(cin, device="cuda")
    guard = torch.zeros(1, device="cuda", dtype=torch.int32)
    ops.bn_bwd_coef(sums, dev(gamma), dev(beta), dev(mean), dev(rstd), M, dg, db, coef, w=dev(w), wab=wab, bias=bias, guard=guard)
    assert int(guard.item()) == int(np.any(np.abs(gamma) < np.abs(beta) / 16))
    np.testing.assert_allclose(host(db), dbeta, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(host(dg), dgamma, rtol=2e-2, atol=0.5)        # xhat recovered from the bf16 activation
    gd, zd, dd_ = dev(g, bf), dev(z, bf), dev(d, bf)
    dd = torch.empty((M, cin), device="cuda", dtype=bf)
    ops.gemm(gd, wab, dd, b_trans=True, A2=zd, epilogue=ops.EPI_AFFINE, shift=bias)
    G = torch.zeros((cin, 2 * c), device="cuda"); dw = torch.zeros((cin, c), device="cuda")
    ops.gemm(dd_, gd, G, a_trans=True, accumulate=True, B2=zd)
    ops.bn_bwd_wgrad_combine(G, coef, dev(d.sum(0)), dw)
    scale = np.abs(dd_ref).max()
    assert np.abs(host(dd) - dd_ref).max() <= 2e-2 * scale
    assert np.linalg.norm(host(dw) - dw_ref) <= 2e-2 * np.linalg.norm(dw_ref)


# ------------------------------------------------------------------------------------------------ fused first block
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(2, 16, 32), (3, 21, 45), (1, 64, 64)])
def test_stem_fwd_bwd(dtype, shape):
    n, h, w = shape
    x = RNG.random((n, h, w, 3)).astype(np.float32)
    wd = RNG.standard_normal((3, 3, 3)).astype(np.float32)
    wp = RNG.standard_normal((3, 64)).astype(np.float32)
    dz = RNG.standard_normal((n, h, w, 64)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, 64).astype(np.float32); sh = RNG.standard_normal(64).astype(np.float32)
    xr, dzr = (bf16_round(x), bf16_round(dz)) if dtype == torch.bfloat16 else (x.astype(np.float64), dz.astype(np.float64))
    d = R.dwconv3x3(xr, wd.astype(np.float64))
    z = d.reshape(-1, 3) @ wp.astype(np.float64)
    xd = dev(x, dtype)
    out = torch.empty((n, h, w, 64), device="cuda", dtype=dtype)
    cs = torch.zeros(64, device="cuda", dtype=torch.float64); cq = torch.zeros_like(cs)
    ops.stem_fwd(xd, dev(wd.reshape(9, 3)), dev(wp), out, colsum=cs, colsq=cq)
    got = host(out).reshape(-1, 64)
    np.testing.assert_allclose(got, z, **tol(dtype))
    np.testing.assert_allclose(cs.cpu().numpy(), got.sum(0), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(cq.cpu().numpy(), (got ** 2).sum(0), rtol=1e-4, atol=1e-2)
    buf = torch.zeros((n, h, w, 128), device="cuda", dtype=dtype)
    ops.stem_fwd(xd, dev(wd.reshape(9, 3)), dev(wp), buf[..., 64:], scale=dev(sc), shift=dev(sh), relu=True)
    np.testing.assert_allclose(host(buf)[..., 64:].reshape(-1, 64), np.maximum(z * sc + sh, 0), **tol(dtype))
    assert np.all(host(buf)[..., :64] == 0)
    gwd = torch.zeros((9, 3), device="cuda"); gwp = torch.zeros((3, 64), device="cuda")
    ops.stem_bwd(xd, dev(dz, dtype), dev(wd.reshape(9, 3)), dev(wp), gwd, gwp)
    dd = (dzr.reshape(-1, 64) @ wp.astype(np.float64).T).reshape(n, h, w, 3)
    _, gwd_ref = R.dwconv3x3_bwd(xr, wd.astype(np.float64), dd)
    np.testing.assert_allclose(host(gwp), d.reshape(-1, 3).T @ dzr.reshape(-1, 64), rtol=1e-3, atol=1e-2)
    np.testing.assert_allclose(host(gwd).reshape(3, 3, 3), gwd_ref, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("nhw", [(2, 19, 37), (2, 19, 36), (1, 5, 4), (3, 8, 128)])     # W % 4 == 0: four-pixel depthwise kernel (bf16)
def test_stem_bwd_folded(dtype, nhw):
    """streaming first-block backward: dz = A*g + B*z + K formed in registers == explicit dz fed to the tiled kernel"""
    n, h, w = nhw
    x = RNG.random((n, h, w, 3)).astype(np.float32)
    wd = RNG.standard_normal((3, 3, 3)).astype(np.float32); wp = RNG.standard_normal((3, 64)).astype(np.float32)
    g = RNG.standard_normal((n, h, w, 64)).astype(np.float32); z = RNG.standard_normal((n, h, w, 64)).astype(np.float32)
    coef = np.stack([RNG.uniform(0.5, 1.5, 64), RNG.standard_normal(64) * 0.1, RNG.standard_normal(64) * 0.05]).astype(np.float32)
    xr, gr, zr = (bf16_round(x), bf16_round(g), bf16_round(z)) if dtype == torch.bfloat16 else (x.astype(np.float64), g.astype(np.float64), z.astype(np.float64))
    dz = coef[0] * gr + coef[1] * zr + coef[2]
    d = R.dwconv3x3(xr, wd.astype(np.float64))
    xd = dev(x, dtype)
    out = torch.empty((n, h, w, 64), device="cuda", dtype=dtype)
    cs = torch.zeros(64, device="cuda", dtype=torch.float64); cq = torch.zeros_like(cs)
    d3 = torch.empty((n, h, w, 3), device="cuda")
    ops.stem_fwd(xd, dev(wd.reshape(9, 3)), dev(wp), out, colsum=cs, colsq=cq, d_out=d3)
    np.testing.assert_allclose(host(d3), d, rtol=1e-5, atol=1e-5)
    # the streaming pair (depthwise into d3, barrier-free pointwise) against the tiled kernel: statistics and BN+ReLU modes
    out_t = torch.empty_like(out); cs_t = torch.zeros_like(cs); cq_t = torch.zeros_like(cq)
    ops.stem_fwd(xd, dev(wd.reshape(9, 3)), dev(wp), out_t, colsum=cs_t, colsq=cq_t)
    np.testing.assert_allclose(host(out), host(out_t), **tol(dtype))
    np.testing.assert_allclose(cs.cpu().numpy(), cs_t.cpu().numpy(), rtol=1e-3, atol=5e-1)
    np.testing.assert_allclose(cq.cpu().numpy(), cq_t.cpu().numpy(), rtol=1e-3, atol=5e-1)
    sc = RNG.uniform(0.5, 1.5, 64).astype(np.float32); sh = RNG.standard_normal(64).astype(np.float32)
    buf_s = torch.zeros((n, h, w, 128), device="cuda", dtype=dtype); buf_t = torch.zeros_like(buf_s)
    ops.stem_fwd(xd, dev(wd.reshape(9, 3)), dev(wp), buf_s[..., 64:], scale=dev(sc), shift=dev(sh), relu=True, d_out=d3)
    ops.stem_fwd(xd, dev(wd.reshape(9, 3)), dev(wp), buf_t[..., 64:], scale=dev(sc), shift=dev(sh), relu=True)
    np.testing.assert_allclose(host(buf_s), host(buf_t), **tol(dtype))
    gwp = torch.zeros((3, 64), device="cuda"); dd = torch.empty((n, h, w, 3), device="cuda", dtype=dtype)
    ops.stem_bwd_folded(dev(g, dtype), dev(z, dtype), dev(coef), d3, dev(wp), gwp, dd)
    np.testing.assert_allclose(host(gwp), d.reshape(-1, 3).T @ dz.reshape(-1, 64), rtol=1e-3, atol=1e-2)
    dd_ref = (dz.reshape(-1, 64) @ wp.astype(np.float64).T).reshape(n, h, w, 3)
    np.testing.assert_allclose(host(dd), dd_ref, **(dict(rtol=1e-4, atol=1e-4) if dtype == torch.float32 else dict(rtol=1e-2, atol=5e-2)))
    gwd = torch.zeros((9, 3), device="cuda")
    ops.dwconv3x3_bwd_weight(xd, dd, gwd)
    _, gwd_ref = R.dwconv3x3_bwd(xr, wd.astype(np.float64), host(dd))
    np.testing.assert_allclose(host(gwd).reshape(3, 3, 3), gwd_ref, rtol=1e-3, atol=1e-2)


@pytest.mark.parametrize("nhw", [(2, 6, 8), (1, 9, 36), (3, 4, 4)])
def test_stem_depthwise_reads_stay_inside_the_image(nhw):
    """the four-pixel depthwise kernel loads aligned 32-bit words around each pixel group: with the image embedded between
    NaN-filled neighbours, any word taken from outside its rows would poison the result"""
    n, h, w = nhw
    big = torch.full((n + 2, h, w, 3), float("nan"), device="cuda", dtype=torch.bfloat16)
    x = torch.rand((n, h, w, 3), device="cuda").to(torch.bfloat16)
    big[1:n + 1] = x
    wd = dev(RNG.standard_normal((9, 3))); wp = dev(RNG.standard_normal((3, 64)))
    out_a = torch.empty((n, h, w, 64), device="cuda", dtype=torch.bfloat16); out_b = torch.empty_like(out_a)
    d_a = torch.full((n + 2, h, w, 3), 7.0, device="cuda"); d_b = torch.empty((n, h, w, 3), device="cuda")
    ops.stem_fwd(big[1:n + 1], wd, wp, out_a, d_out=d_a[1:n + 1])
    ops.stem_fwd(x.clone(), wd, wp, out_b, d_out=d_b)
    assert bool(torch.isfinite(d_a[1:n + 1]).all()) and torch.equal(d_a[1:n + 1], d_b) and torch.equal(out_a, out_b)
    assert bool((d_a[0] == 7.0).all()) and bool((d_a[n + 1] == 7.0).all())          # and nothing is written outside either
    ref = R.dwconv3x3(host(x).astype(np.float64), host(wd).reshape(3, 3, 3).astype(np.float64))
    np.testing.assert_allclose(host(d_b), ref, rtol=1e-5, atol=1e-5)


def test_cast_transpose_bf16_batched_matches_per_matrix():
    """every dense kernel of the model staged by one launch == one launch per matrix (ragged sizes, either destination absent)"""
    shapes = [(64, 64), (3, 64), (70, 33), (256, 128), (1, 5), (128, 1024)]
    base = torch.randn(sum(r * c for r, c in shapes) + 7, device="cuda")
    items, refs, off = [], [], 3
    for k, (r, c) in enumerate(shapes):
        dst = None if k == 2 else torch.full((r, c), 9.0, device="cuda", dtype=torch.bfloat16)
        dst_t = None if k == 4 else torch.full((c, r), 9.0, device="cuda", dtype=torch.bfloat16)
        items.append((off, dst, dst_t, r, c))
        a = torch.empty((r, c), device="cuda", dtype=torch.bfloat16); at = torch.empty((c, r), device="cuda", dtype=torch.bfloat16)
        ops.cast_transpose_bf16(base[off:off + r * c].view(r, c), a, at)
        refs.append((a, at))
        off += r * c
    table, n, tiles = ops.cast_transpose_table(base, items)
    ops.cast_transpose_bf16_batched(base, table, n, tiles)
    for (o, dst, dst_t, r, c), (a, at) in zip(items, refs):
        assert dst is None or torch.equal(dst, a)
        assert dst_t is None or torch.equal(dst_t, at)
        assert torch.equal(a, base[o:o + r * c].view(r, c).to(torch.bfloat16))
    with pytest.raises(ValueError):
        ops.cast_transpose_table(base, [(base.numel() - 3, None, None, 2, 2)])


# ------------------------------------------------------------------------------------------------ GEMM, CUDA cores
@pytest.mark.parametrize("a_trans,b_trans", [(False, False), (False, True), (True, False)])
@pytest.mark.parametrize("mkn", [(70, 3, 64), (129, 40, 33), (64, 64, 1), (256, 128, 96)])
def test_gemm_simt_fp32(a_trans, b_trans, mkn):
    M, K, N = mkn
    A = RNG.standard_normal((K, M) if a_trans else (M, K)).astype(np.float32)
    B = RNG.standard_normal((N, K) if b_trans else (K, N)).astype(np.float32)
    ref = (A.T if a_trans else A).astype(np.float64) @ (B.T if b_trans else B).astype(np.float64)
    Cm = torch.zeros((M, N), device="cuda")
    ops.gemm(dev(A), dev(B), Cm, a_trans=a_trans, b_trans=b_trans, accumulate=a_trans)
    np.testing.assert_allclose(host(Cm), ref, rtol=1e-5, atol=1e-4)


def test_gemm_simt_epilogues():
    M, K, N = 300, 24, 40
    A = RNG.standard_normal((M, K)).astype(np.float32)
    B = RNG.standard_normal((K, N)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, N).astype(np.float32)
    sh = RNG.standard_normal(N).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    Cm = torch.empty((M, N), device="cuda")
    ops.gemm(dev(A), dev(B), Cm, epilogue=ops.EPI_AFFINE_RELU, scale=dev(sc), shift=dev(sh))
    np.testing.assert_allclose(host(Cm), np.maximum(ref * sc + sh, 0), rtol=1e-5, atol=1e-4)
    cs = torch.zeros(N, device="cuda", dtype=torch.float64)
    cq = torch.zeros(N, device="cuda", dtype=torch.float64)
    ops.gemm(dev(A), dev(B), Cm, epilogue=ops.EPI_STATS, colsum=cs, colsq=cq)
    np.testing.assert_allclose(host(Cm), ref, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(cs.cpu().numpy(), ref.sum(0), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(cq.cpu().numpy(), (ref ** 2).sum(0), rtol=1e-5, atol=1e-3)


def _convt_case(dtype, n, h, w, cin, cout, rate):
    x = RNG.standard_normal((n, h, w, cin)).astype(np.float32)
    k = (RNG.standard_normal((2, 2, cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = RNG.standard_normal(cout).astype(np.float32)
    xr, kr = (bf16_round(x), bf16_round(k)) if dtype == torch.bfloat16 else (x.astype(np.float64), k.astype(np.float64))
    ref = R.convt2x2(xr, kr, b.astype(np.float64))
    Bkn = k.transpose(3, 0, 1, 2).reshape(cin, 4 * cout)          # [K, (a,b,co)]
    concat = torch.zeros((n, 2 * h, 2 * w, 2 * cout), device="cuda", dtype=dtype)
    drop = ops.make_dropout(rate, 5, ctot=2 * cout, c0=0)
    if dtype == torch.bfloat16:
        ops.gemm(dev(x, dtype), dev(Bkn.T.copy(), dtype), concat[..., :cout], b_trans=True, epilogue=ops.EPI_CONVT,
                 shift=dev(b), convt_hw=(h, w), drop=drop)
    else:
        ops.gemm(dev(x), dev(Bkn), concat[..., :cout], epilogue=ops.EPI_CONVT, shift=dev(b), convt_hw=(h, w), drop=drop)
    if rate > 0:
        ref = ref * R.dropout_multiplier((n, 2 * h, 2 * w, 2 * cout), rate, 5)[..., :cout]
    got = host(concat)
    np.testing.assert_allclose(got[..., :cout], ref, **tol(dtype))
    assert np.all(got[..., cout:] == 0)


@pytest.mark.parametrize("rate", [0.0, 0.2])
def test_convt_simt(rate):
    _convt_case(torch.float32, 2, 3, 5, 24, 12, rate)


# ------------------------------------------------------------------------------------------------ GEMM, tcgen05
TC_SHAPES = [(128, 64, 64), (1000, 64, 64), (300, 72, 128), (4096, 512, 512), (257, 1024, 256), (640, 128, 1024),
             (129, 8, 64), (2048, 256, 192), (96, 1024, 2048)]


@pytest.mark.parametrize("mkn", TC_SHAPES)
@pytest.mark.parametrize("out_dtype", DTYPES)
def test_gemm_tc_nt(mkn, out_dtype):
    M, K, N = mkn
    A = RNG.standard_normal((M, K)).astype(np.float32)
    Bt = (RNG.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    ref = bf16_round(A) @ bf16_round(Bt).T
    Cm = torch.full((M, N), float("nan"), device="cuda", dtype=out_dtype)
    ops.gemm(dev(A, torch.bfloat16), dev(Bt, torch.bfloat16), Cm, b_trans=True, tensor_core=True)
    np.testing.assert_allclose(host(Cm), ref, **(dict(rtol=1e-4, atol=1e-4) if out_dtype == torch.float32 else tol(out_dtype)))


def test_gemm_tc_nt_strided_affine_relu():
    M, K, N = 900, 128, 128
    Abuf = RNG.standard_normal((M, 256)).astype(np.float32)
    Bt = (RNG.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, N).astype(np.float32)
    sh = RNG.standard_normal(N).astype(np.float32)
    Ad = dev(Abuf, torch.bfloat16)
    out = torch.zeros((M, 384), device="cuda", dtype=torch.bfloat16)
    ops.gemm(Ad[:, 64:64 + K], dev(Bt, torch.bfloat16), out[:, 128:128 + N], b_trans=True, tensor_core=True,
             epilogue=ops.EPI_AFFINE_RELU, scale=dev(sc), shift=dev(sh))
    ref = np.maximum((bf16_round(Abuf)[:, 64:64 + K] @ bf16_round(Bt).T) * sc + sh, 0)
    got = host(out)
    np.testing.assert_allclose(got[:, 128:128 + N], ref, **tol(torch.bfloat16))
    assert np.all(got[:, :128] == 0) and np.all(got[:, 256:] == 0)


@pytest.mark.parametrize("classes", [1, 3, 8])
@pytest.mark.parametrize("store_y", [False, True])
def test_gemm_tc_fused_head(classes, store_y):
    """dec1_block2 pointwise + folded BN + ReLU + Conv2D(classes, 1, sigmoid|softmax) in one epilogue (u_net.py:105-112)."""
    M, K, N = 1000, 64, 64
    A = RNG.standard_normal((M, K)).astype(np.float32)
    Bt = (RNG.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, N).astype(np.float32); sh = RNG.standard_normal(N).astype(np.float32) * 0.3
    hw = (RNG.standard_normal((N, classes)) / 4).astype(np.float32); hb = RNG.standard_normal(classes).astype(np.float32) * 0.1
    y_exact = np.maximum((bf16_round(A) @ bf16_round(Bt).T) * sc + sh, 0)
    y = bf16_round(y_exact)
    logits = y_exact @ hw.astype(np.float64) + hb            # the head reads the fp32 activations in registers, not the stored bf16
    ref = R.sigmoid(logits) if classes == 1 else R.softmax(logits)
    Cm = torch.zeros((M, N), device="cuda", dtype=torch.bfloat16) if store_y else None
    probs = torch.full((M, classes), float("nan"), device="cuda")
    ops.gemm(dev(A, torch.bfloat16), dev(Bt, torch.bfloat16), Cm, b_trans=True, epilogue=ops.EPI_HEAD, scale=dev(sc), shift=dev(sh),
             head_w=dev(hw), head_b=dev(hb), head_out=probs)
    np.testing.assert_allclose(host(probs), ref, rtol=0, atol=3e-3)
    if store_y:
        np.testing.assert_allclose(host(Cm), y, **tol(torch.bfloat16))


@pytest.mark.parametrize("mkn", [(128, 32, 64), (1000, 64, 64), (300, 72, 128), (4096, 512, 512), (257, 1024, 256), (129, 8, 64), (96, 36, 1024)])
def test_gemm_tc_fp32_tf32x3(mkn):
    """fp32 mode on the tensor cores: operands split into tf32 (hi, lo) parts, three kind::tf32 MMAs per k-step.  Against fp64:
    the error must be fp32-grade (a single tf32 product would be off by ~5e-4 relative), incl. strided A views and ragged tiles."""
    M, K, N = mkn
    a = RNG.standard_normal((M, K)).astype(np.float32)
    b = (RNG.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64).T
    abuf = torch.zeros((M, K + 8), device="cuda"); abuf[:, 4:4 + K] = dev(a)
    bh, bl = dev(b), torch.empty((N, K), device="cuda")
    ops.split_tf32(bh, None, bl)
    trunc = (bh.view(torch.int32) & ~0x1FFF).view(torch.float32)                      # what kind::tf32 multiplies when it reads b
    assert float((trunc.double() + bl.double() - bh.double()).abs().max()) <= 2.0 ** -21 * float(bh.abs().max())
    assert bool(((bl.view(torch.int32) & 0x1FFF) == 0).all())                          # lo is a tf32 value
    c = torch.zeros((M, N + 4), device="cuda")
    ops.gemm(abuf[:, 4:4 + K], bh, c[:, :N], b_trans=True, B_lo=bl, tensor_core=True)
    got = host(c)
    err = np.abs(got[:, :N] - ref).max()
    assert err <= 2e-5 * np.abs(ref).max() + 1e-6, err
    assert np.all(got[:, N:] == 0)
    # transposed split (weights kept as [K, N] in Keras order)
    bth, btl = torch.empty((N, K), device="cuda"), torch.empty((N, K), device="cuda")
    ops.split_tf32(dev(b.T.copy()), bth, btl, transpose=True)
    np.testing.assert_array_equal(host(bth), host(bh)); np.testing.assert_array_equal(host(btl), host(bl))       # hi = raw copy
    # epilogues: scale/shift + ReLU, batch statistics
    sc, sh = RNG.uniform(0.5, 1.5, N).astype(np.float32), RNG.standard_normal(N).astype(np.float32)
    c2 = torch.empty((M, N), device="cuda")
    ops.gemm(dev(a), bh, c2, b_trans=True, B_lo=bl, epilogue=ops.EPI_AFFINE_RELU, scale=dev(sc), shift=dev(sh))
    np.testing.assert_allclose(host(c2), np.maximum(ref * sc + sh, 0), rtol=1e-5, atol=2e-5 * np.abs(ref).max())
    cs, cq = torch.zeros(N, device="cuda", dtype=torch.float64), torch.zeros(N, device="cuda", dtype=torch.float64)
    ops.gemm(dev(a), bh, c2, b_trans=True, B_lo=bl, epilogue=ops.EPI_STATS, colsum=cs, colsq=cq)
    np.testing.assert_allclose(cs.cpu().numpy(), ref.sum(0), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(cq.cpu().numpy(), (ref ** 2).sum(0), rtol=2e-5, atol=1e-3)


@pytest.mark.parametrize("pmn", [(4096, 64, 64), (10000, 128, 256), (777, 1024, 512), (20000, 8, 64), (3000, 512, 1024), (50, 36, 72)])
def test_gemm_tc_wgrad_fp32_tf32x3(pmn):
    """fp32 weight gradient C[M,N] += A[P,M]^T B[P,N] on the tensor cores (both operands MN-major, tf32 split of both)."""
    P, M, N = pmn
    a = RNG.standard_normal((P, M)).astype(np.float32)
    b = RNG.standard_normal((P, N)).astype(np.float32)
    ref = a.astype(np.float64).T @ b.astype(np.float64)
    abuf = torch.zeros((P, M + 8), device="cuda"); abuf[:, 4:4 + M] = dev(a)
    c0 = RNG.standard_normal((M, N)).astype(np.float32)
    c = dev(c0)
    ops.gemm(abuf[:, 4:4 + M], dev(b), c, a_trans=True, accumulate=True, tf32x3=True, tensor_core=True)
    got = host(c) - c0
    assert np.abs(got - ref).max() <= 3e-5 * np.sqrt(P) + 1e-5, np.abs(got - ref).max()     # |C| ~ 4 sqrt(P): ~7e-6 relative, fp32 atomics included
    # against the CUDA-core kernel on the same inputs
    c2 = torch.zeros((M, N), device="cuda")
    ops.gemm(dev(a), dev(b), c2, a_trans=True, accumulate=True)
    np.testing.assert_allclose(got, host(c2), rtol=0, atol=3e-5 * np.sqrt(P) + 1e-5)


@pytest.mark.parametrize("rate", [0.0, 0.2])
@pytest.mark.parametrize("cfg", [(2, 4, 4, 128, 64), (1, 8, 16, 256, 32), (3, 6, 128, 64, 64)])
def test_convt_tc_fp32_tf32x3(cfg, rate):
    n, h, w, cin, cout = cfg
    x = RNG.standard_normal((n, h, w, cin)).astype(np.float32)
    k = (RNG.standard_normal((2, 2, cout, cin)) / np.sqrt(cin)).astype(np.float32)
    b = RNG.standard_normal(cout).astype(np.float32)
    ref = R.convt2x2(x.astype(np.float64), k.astype(np.float64), b.astype(np.float64))
    bh = dev(k.reshape(4 * cout, cin))
    bl = torch.empty_like(bh)
    ops.split_tf32(bh, None, bl)
    concat = torch.zeros((n, 2 * h, 2 * w, 2 * cout), device="cuda")
    drop = ops.make_dropout(rate, 5, ctot=2 * cout, c0=0)
    ops.gemm(dev(x), bh, concat[..., :cout], b_trans=True, B_lo=bl, epilogue=ops.EPI_CONVT, shift=dev(b), convt_hw=(h, w), drop=drop,
             tensor_core=True)
    if rate > 0:
        ref = ref * R.dropout_multiplier((n, 2 * h, 2 * w, 2 * cout), rate, 5)[..., :cout]
    got = host(concat)
    np.testing.assert_allclose(got[..., :cout], ref, rtol=2e-5, atol=3e-5)
    assert np.all(got[..., cout:] == 0)


@pytest.mark.parametrize("mkn", [(1000, 64, 64), (5000, 256, 512), (333, 128, 192)])
@pytest.mark.parametrize("out_dtype", DTYPES)
def test_gemm_tc_stats(mkn, out_dtype):
    M, K, N = mkn
    A = RNG.standard_normal((M, K)).astype(np.float32)
    Bt = (RNG.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    Cm = torch.empty((M, N), device="cuda", dtype=out_dtype)
    cs = torch.zeros(N, device="cuda", dtype=torch.float64)
    cq = torch.zeros(N, device="cuda", dtype=torch.float64)
    ops.gemm(dev(A, torch.bfloat16), dev(Bt, torch.bfloat16), Cm, b_trans=True, tensor_core=True,
             epilogue=ops.EPI_STATS, colsum=cs, colsq=cq)
    got = host(Cm)     # the statistics are defined over the STORED values
    np.testing.assert_allclose(got, bf16_round(A) @ bf16_round(Bt).T, **(dict(rtol=1e-4, atol=1e-4) if out_dtype == torch.float32 else tol(out_dtype)))
    np.testing.assert_allclose(cs.cpu().numpy(), got.sum(0), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(cq.cpu().numpy(), (got ** 2).sum(0), rtol=1e-4, atol=1e-2)


@pytest.mark.parametrize("rate", [0.0, 0.2])
@pytest.mark.parametrize("cfg", [(2, 4, 4, 128, 64), (1, 8, 6, 1024, 512), (3, 5, 3, 256, 128), (3, 6, 16, 128, 64),
                                 (1, 3, 256, 64, 64), (2, 20, 128, 72, 128)])
def test_convt_tc(cfg, rate):
    _convt_case(torch.bfloat16, *cfg, rate)


@pytest.mark.parametrize("pmn", [(4096, 64, 64), (10000, 128, 256), (777, 1024, 512), (20000, 8, 64), (3000, 512, 1024)])
def test_gemm_tc_wgrad(pmn):
    P, Mo, No = pmn
    A = RNG.standard_normal((P, Mo)).astype(np.float32)
    B = RNG.standard_normal((P, No)).astype(np.float32)
    ref = bf16_round(A).T @ bf16_round(B)
    Cm = torch.zeros((Mo, No), device="cuda")
    ops.gemm(dev(A, torch.bfloat16), dev(B, torch.bfloat16), Cm, a_trans=True, accumulate=True, tensor_core=True)
    np.testing.assert_allclose(host(Cm), ref, rtol=1e-4, atol=1e-3 * np.sqrt(P))


# ------------------------------------------------------------------------------------------------ batch norm
@pytest.mark.parametrize("dtype", DTYPES)
def test_bn_train_fwd_bwd(dtype):
    n, h, w, c = 2, 8, 12, 64
    z = RNG.standard_normal((n, h, w, c)).astype(np.float32) * 2 + 0.5
    dy = RNG.standard_normal((n, h, w, c)).astype(np.float32)
    gamma = RNG.uniform(0.5, 1.5, c).astype(np.float32)
    beta = RNG.standard_normal(c).astype(np.float32) * 0.2
    mm = RNG.standard_normal(c).astype(np.float32)
    mv = RNG.uniform(0.5, 1.5, c).astype(np.float32)
    zr, dyr = (bf16_round(z), bf16_round(dy)) if dtype == torch.bfloat16 else (z.astype(np.float64), dy.astype(np.float64))
    M = n * h * w
    mean, var = zr.mean((0, 1, 2)), zr.var((0, 1, 2))
    rstd = 1 / np.sqrt(var + 1e-3)
    xhat = (zr - mean) * rstd
    y_ref = np.maximum(xhat * gamma + beta, 0)

    cs = dev(zr.reshape(M, c).sum(0), torch.float64)
    cq = dev((zr.reshape(M, c) ** 2).sum(0), torch.float64)
    f = lambda: torch.empty(c, device="cuda")
    scale, shift, smean, srstd = f(), f(), f(), f()
    mmd, mvd = dev(mm), dev(mv)
    ops.bn_finalize(cs, cq, M, dev(gamma), dev(beta), 1e-3, 0.99, mmd, mvd, scale, shift, smean, srstd)
    np.testing.assert_allclose(host(smean), mean, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(host(srstd), rstd, rtol=1e-5)
    np.testing.assert_allclose(host(mmd), mm * 0.99 + mean * 0.01, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(host(mvd), mv * 0.99 + var * 0.01, rtol=1e-5, atol=1e-6)

    zd = dev(z, dtype)
    y = torch.empty_like(zd)
    pooled = torch.empty((n, h // 2, w // 2, c), device="cuda", dtype=dtype)
    ops.bn_act(zd, scale, shift, y, relu=True, pooled=pooled)
    np.testing.assert_allclose(host(y), y_ref, **tol(dtype))
    np.testing.assert_allclose(host(pooled), R.maxpool2x2(host(y)), rtol=0, atol=0)

    g = dyr * (y_ref > 0)
    dgamma_ref, dbeta_ref = (g * xhat).sum((0, 1, 2)), g.sum((0, 1, 2))
    dz_ref = gamma * rstd * (g - dbeta_ref / M - xhat * dgamma_ref / M)
    dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    dyd = dev(dy, dtype)
    ops.bn_bwd_reduce(dyd, zd, scale, shift, smean, srstd, dg, db)
    np.testing.assert_allclose(host(dg), dgamma_ref, rtol=1e-3, atol=2e-2)
    np.testing.assert_allclose(host(db), dbeta_ref, rtol=1e-3, atol=2e-2)
    dz = torch.empty_like(zd)
    ops.bn_bwd_apply(dyd, zd, scale, shift, smean, srstd, dg, db, dz)
    np.testing.assert_allclose(host(dz), dz_ref, **(dict(rtol=1e-3, atol=1e-4) if dtype == torch.float32 else tol(dtype)))


def test_bn_fold_and_act_dropout():
    c = 32
    gamma, beta = RNG.uniform(0.5, 1.5, c).astype(np.float32), RNG.standard_normal(c).astype(np.float32)
    mean, var = RNG.standard_normal(c).astype(np.float32), RNG.uniform(0.5, 1.5, c).astype(np.float32)
    scale, shift = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    ops.bn_fold(dev(gamma), dev(beta), dev(mean), dev(var), 1e-3, scale, shift)
    s_ref = gamma / np.sqrt(var + 1e-3)
    np.testing.assert_allclose(host(scale), s_ref, rtol=1e-6)
    np.testing.assert_allclose(host(shift), beta - mean * s_ref, rtol=1e-5, atol=1e-6)
    z = RNG.standard_normal((2, 4, 6, c)).astype(np.float32)
    buf = torch.zeros((2, 4, 6, 2 * c), device="cuda")
    drop = ops.make_dropout(0.2, 9, ctot=2 * c, c0=c)
    ops.bn_act(dev(z), scale, shift, buf[..., c:], relu=True, drop=drop)
    ref = np.maximum(z * host(scale) + host(shift), 0) * R.dropout_multiplier((2, 4, 6, 2 * c), 0.2, 9)[..., c:]
    np.testing.assert_allclose(host(buf)[..., c:], ref, rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------ pooling
@pytest.mark.parametrize("dtype", DTYPES)
def test_maxpool_fwd_bwd(dtype):
    n, h, w, c = 2, 6, 10, 16
    x = np.maximum(RNG.standard_normal((n, h, w, c)), 0).astype(np.float32)   # post-ReLU: many exact ties at 0
    xr = bf16_round(x) if dtype == torch.bfloat16 else x.astype(np.float64)
    buf = torch.zeros((n, h, w, 3 * c), device="cuda", dtype=dtype)
    buf[..., c:2 * c] = dev(x, dtype)
    y = torch.empty((n, h // 2, w // 2, c), device="cuda", dtype=dtype)
    ops.maxpool2x2(buf[..., c:2 * c], y)
    np.testing.assert_array_equal(host(y), R.maxpool2x2(xr))
    dpool = RNG.standard_normal((n, h // 2, w // 2, c)).astype(np.float32)
    dskip = RNG.standard_normal((n, h, w, c)).astype(np.float32)
    dpr, dsr = (bf16_round(dpool), bf16_round(dskip)) if dtype == torch.bfloat16 else (dpool.astype(np.float64), dskip.astype(np.float64))
    dy = torch.empty((n, h, w, c), device="cuda", dtype=dtype)
    ops.maxpool2x2_bwd(buf[..., c:2 * c], None, None, dev(dpool, dtype), dev(dskip, dtype), dy)
    np.testing.assert_allclose(host(dy), R.maxpool2x2_bwd(xr, dpr) + dsr, **tol(dtype))
    # activation recomputed from z, ReLU-masked output and BatchNormalization's backward reductions
    z = RNG.standard_normal((n, h, w, c)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, c).astype(np.float32); sh = (RNG.standard_normal(c) * 0.3).astype(np.float32)
    zd = dev(z, dtype)
    yact = torch.empty_like(zd)
    ops.bn_act(zd, dev(sc), dev(sh), yact, relu=True)
    ya = host(yact)
    plain, masked = torch.empty_like(zd), torch.empty_like(zd)
    bsum = torch.zeros((2, c), device="cuda")
    ops.maxpool2x2_bwd(zd, dev(sc), dev(sh), dev(dpool, dtype), dev(dskip, dtype), plain)
    ops.maxpool2x2_bwd(zd, dev(sc), dev(sh), dev(dpool, dtype), dev(dskip, dtype), masked, bn_sums=bsum)
    np.testing.assert_allclose(host(plain), R.maxpool2x2_bwd(ya, dpr) + dsr, **tol(dtype))
    dropped = torch.empty_like(zd)
    cat = torch.zeros((n, h, w, 2 * c), device="cuda", dtype=dtype); cat[..., c:] = dev(dskip, dtype)
    ops.maxpool2x2_bwd(zd, dev(sc), dev(sh), dev(dpool, dtype), cat[..., c:], dropped, skip_drop=ops.make_dropout(0.25, 23, ctot=2 * c, c0=c))
    mult = R.dropout_multiplier((n, h, w, 2 * c), 0.25, 23)[..., c:]
    np.testing.assert_allclose(host(dropped), R.maxpool2x2_bwd(ya, dpr) + dsr * mult, **tol(dtype))
    gm = host(masked)
    np.testing.assert_array_equal(gm, host(plain) * (ya > 0))
    np.testing.assert_allclose(host(bsum)[0], gm.sum((0, 1, 2)), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(host(bsum)[1], (gm * ya).sum((0, 1, 2)), rtol=1e-4, atol=1e-4)


def test_convt_bwd_gather():
    n, h, w, co = 2, 3, 4, 16
    du = RNG.standard_normal((n, 2 * h, 2 * w, co)).astype(np.float32)
    buf = torch.zeros((n, 2 * h, 2 * w, 2 * co), device="cuda")
    buf[..., :co] = dev(du)
    g = torch.empty((n * h * w, 4 * co), device="cuda")
    db = torch.zeros(co, device="cuda")
    ops.convt_bwd_gather(buf[..., :co], g, db)
    ref = du.reshape(n, h, 2, w, 2, co).transpose(0, 1, 3, 2, 4, 5).reshape(n * h * w, 4 * co)
    np.testing.assert_array_equal(host(g), ref)
    np.testing.assert_allclose(host(db), du.sum((0, 1, 2)), rtol=1e-5, atol=1e-5)
    # the concat buffer's Dropout mask applied on the way (deferred from the kernel that wrote du)
    g2 = torch.empty_like(g); db2 = torch.zeros(co, device="cuda")
    ops.convt_bwd_gather(buf[..., :co], g2, db2, drop=ops.make_dropout(0.25, 19, ctot=2 * co, c0=0))
    dud = du * R.dropout_multiplier((n, 2 * h, 2 * w, 2 * co), 0.25, 19)[..., :co]
    np.testing.assert_allclose(host(g2), dud.reshape(n, h, 2, w, 2, co).transpose(0, 1, 3, 2, 4, 5).reshape(n * h * w, 4 * co), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(host(db2), dud.sum((0, 1, 2)), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------ head + loss
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("C", [1, 8])
@pytest.mark.parametrize("kind", [0, 1])
def test_head_and_loss(dtype, C, kind):
    n, h, w, k = 3, 10, 12, 64
    x = RNG.standard_normal((n, h, w, k)).astype(np.float32)
    wk = (RNG.standard_normal((k, C)) / 4).astype(np.float32)
    b = RNG.standard_normal(C).astype(np.float32) * 0.1
    _, t = R.synthetic_batch(n, h, w, 3, C, seed=5)
    xr = bf16_round(x) if dtype == torch.bfloat16 else x.astype(np.float64)
    logits = xr.reshape(-1, k) @ wk.astype(np.float64) + b
    p_ref = (R.sigmoid(logits) if C == 1 else R.softmax(logits)).reshape(n, h, w, C)
    xd = dev(x, dtype)
    probs = torch.empty((n, h, w, C), device="cuda")
    sums = torch.zeros((n, C, 3), device="cuda", dtype=torch.float64)
    td = dev(t)
    ops.head_fwd(xd, dev(wk), dev(b), probs, td, sums)
    np.testing.assert_allclose(host(probs), p_ref, rtol=1e-5, atol=1e-6)
    pg = host(probs)
    np.testing.assert_allclose(sums.cpu().numpy()[..., 0], (t * pg).sum((1, 2)), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(sums.cpu().numpy()[..., 1], t.sum((1, 2)), rtol=1e-6)
    np.testing.assert_allclose(sums.cpu().numpy()[..., 2], pg.sum((1, 2)), rtol=1e-5)
    out3 = torch.empty(3, device="cuda")
    coef = torch.empty((n, C, 2), device="cuda")
    ops.seg_loss_finalize(sums, n * C, R.EPSILON, kind, 1.0, out3, coef)
    o = host(out3)
    d_ref, i_ref = R.dice_coef(t, pg, dtype=np.float64), R.iou_coef(t, pg, dtype=np.float64)
    np.testing.assert_allclose(o, [1 - (d_ref if kind == 0 else i_ref), d_ref, i_ref], rtol=1e-5, atol=1e-6)
    # backward through the activation and the 1x1 convolution vs analytic formula
    I, T, P = (t * p_ref).sum((1, 2)), t.sum((1, 2)), p_ref.sum((1, 2))
    if kind == 0:
        D = T + P + R.EPSILON
        dp = -(1.0 / (n * C)) * (2 * t * D[:, None, None] - (2 * I + R.EPSILON)[:, None, None]) / (D ** 2)[:, None, None]
    else:
        U = T + P - I + R.EPSILON
        dp = -(1.0 / (n * C)) * (t * U[:, None, None] - (I + R.EPSILON)[:, None, None] * (1 - t)) / (U ** 2)[:, None, None]
    dl = dp * p_ref * (1 - p_ref) if C == 1 else p_ref * (dp - (dp * p_ref).sum(-1, keepdims=True))
    dl2 = dl.reshape(-1, C)
    dx = torch.empty_like(xd)
    dw, db = torch.zeros((k, C), device="cuda"), torch.zeros(C, device="cuda")
    ops.head_bwd(xd, dev(wk), probs, td, coef, dx, dw, db)
    np.testing.assert_allclose(host(dw), xr.reshape(-1, k).T @ dl2, rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(host(db), dl2.sum(0), rtol=2e-3, atol=2e-6)
    ref_dx = (dl2 @ wk.astype(np.float64).T).reshape(n, h, w, k)
    np.testing.assert_allclose(host(dx), ref_dx, rtol=1e-2 if dtype == torch.bfloat16 else 1e-3, atol=1e-7)
    # ReLU mask of the producing block + BatchNormalization's backward reductions from the same pass (x is then a
    # post-ReLU activation: x >= 0, which the kernel relies on for sum(g*x))
    xp = np.maximum(x, 0); xpr = np.maximum(xr, 0)
    xpd = dev(xp, dtype)
    dxp, dxm = torch.empty_like(xd), torch.empty_like(xd); bsum = torch.zeros((2, k), device="cuda")
    dw1, db1 = torch.zeros((k, C), device="cuda"), torch.zeros(C, device="cuda")
    dw2, db2 = torch.zeros((k, C), device="cuda"), torch.zeros(C, device="cuda")
    ops.head_bwd(xpd, dev(wk), probs, td, coef, dxp, dw1, db1)
    ops.head_bwd(xpd, dev(wk), probs, td, coef, dxm, dw2, db2, bn_sums=bsum)
    gm = host(dxm)
    np.testing.assert_allclose(gm, host(dxp) * (xpr > 0), rtol=0, atol=0)
    np.testing.assert_allclose(host(dw2), host(dw1), rtol=1e-4, atol=1e-7)
    np.testing.assert_allclose(host(bsum)[0], gm.sum((0, 1, 2)), rtol=1e-2 if dtype == torch.bfloat16 else 1e-4, atol=1e-5)
    np.testing.assert_allclose(host(bsum)[1], (gm * xpr).sum((0, 1, 2)), rtol=1e-2 if dtype == torch.bfloat16 else 1e-4, atol=1e-5)


def test_head_stream_affine_on_load():
    """streamed binary head with dec1_block2's BN+ReLU applied on load == the same kernels on the materialised activation"""
    n, h, w, k = 3, 20, 24, 64
    bf = torch.bfloat16
    z = RNG.standard_normal((n, h, w, k)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, k).astype(np.float32); sh = (RNG.standard_normal(k) * 0.4).astype(np.float32)
    wk = (RNG.standard_normal((k, 1)) / 4).astype(np.float32); b = np.array([0.1], np.float32)
    _, t = R.synthetic_batch(n, h, w, 3, 1, seed=7)
    zd = dev(z, bf)
    yd = torch.clamp_min(zd.float() * dev(sc) + dev(sh), 0)          # fp32 activation the kernels form in registers
    assert ops.head_stream_supported(zd, 1)
    res = []
    for x_in, aff in ((zd, True), (yd.to(bf), False)):
        probs = torch.empty((n, h, w, 1), device="cuda"); sums = torch.zeros((n, 1, 3), device="cuda", dtype=torch.float64)
        ops.head_fwd(x_in, dev(wk), dev(b), probs, dev(t), sums, x_scale=dev(sc) if aff else None, x_shift=dev(sh) if aff else None)
        out3 = torch.empty(3, device="cuda"); coef = torch.empty((n, 1, 2), device="cuda")
        ops.seg_loss_finalize(sums, n, R.EPSILON, 0, 1.0, out3, coef)
        dx = torch.empty_like(zd); dw = torch.zeros((k, 1), device="cuda"); db = torch.zeros(1, device="cuda"); bs = torch.zeros((2, k), device="cuda")
        ops.head_bwd(x_in, dev(wk), probs, dev(t), coef, dx, dw, db, bn_sums=bs, x_scale=dev(sc) if aff else None, x_shift=dev(sh) if aff else None)
        res.append((host(probs), host(dx), host(dw), host(db), host(bs)))
    a, m = res          # m rounds the activation to bf16 first: agreement to bf16 noise
    np.testing.assert_allclose(a[0], m[0], atol=5e-3)
    np.testing.assert_allclose(a[2], m[2], rtol=2e-2, atol=1e-4)
    np.testing.assert_allclose(a[3], m[3], rtol=2e-2, atol=1e-5)
    np.testing.assert_allclose(a[4], m[4], rtol=3e-2, atol=1e-4)
    # the mask itself is exact: dx is zero exactly where the activation is zero
    y_exact = host(yd)
    assert np.all(a[1][y_exact <= 0] == 0)


def test_seg_sums_metrics():
    t = (RNG.random((4, 9, 7, 3)) > 0.5).astype(np.float32)
    p = RNG.random((4, 9, 7, 3)).astype(np.float32)
    sums = torch.zeros((4, 3, 3), device="cuda", dtype=torch.float64)
    ops.seg_sums(dev(t), dev(p), sums)
    s = sums.cpu().numpy()
    np.testing.assert_allclose(s[..., 0], (t * p).sum((1, 2)), rtol=1e-5)
    np.testing.assert_allclose(s[..., 1], t.sum((1, 2)), rtol=1e-6)
    np.testing.assert_allclose(s[..., 2], p.sum((1, 2)), rtol=1e-5)


# ------------------------------------------------------------------------------------------------ MeanIoU, AdamW, staging
def test_confusion_matrix():
    n = 100003
    t = RNG.integers(0, 2, n).astype(np.float32)
    p = RNG.random(n).astype(np.float32)
    p[::7] = 1.0
    counts = torch.zeros(4, device="cuda", dtype=torch.int64)
    ops.confusion_matrix_update(dev(t), dev(p), 2, counts)
    m = R.MeanIoU(2); m.update_state(t, p)
    np.testing.assert_array_equal(counts.cpu().numpy().reshape(2, 2), m.cm)
    counts.zero_()
    ops.confusion_matrix_update(dev(t), dev(p), 2, counts, threshold=0.5)
    m = R.MeanIoU(2); m.update_state(t, (p > 0.5).astype(np.uint8))
    np.testing.assert_array_equal(counts.cpu().numpy().reshape(2, 2), m.cm)
    t8 = RNG.integers(0, 8, n).astype(np.float32); p8 = RNG.integers(0, 8, n).astype(np.float32)
    c8 = torch.zeros(64, device="cuda", dtype=torch.int64)
    ops.confusion_matrix_update(dev(t8), dev(p8), 8, c8)
    m = R.MeanIoU(8); m.update_state(t8, p8)
    np.testing.assert_array_equal(c8.cpu().numpy().reshape(8, 8), m.cm)


def test_sample_confusion_thr_matches_calculate_sample_iou():
    """a13: calculate_sample_iou (scripts/benchmark.py:159-170) of thresholded predictions (:260) from per-sample integer counts
    made on the device equals the oracle's float32 restatement bit for bit, and the counts sum to MeanIoU(2)'s matrix (:269)."""
    from unet_b200 import imaging
    rng = np.random.default_rng(3)
    B, h, w = 5, 64, 48
    prob = rng.random((B, h, w, 1)).astype(np.float32)
    prob[3] = 0.1                                                  # empty prediction
    truth = (rng.random((B, h, w, 1)) > 0.6).astype(np.uint8)
    truth[4] = 0                                                   # empty truth
    counts = imaging.gpu_batch_sample_counts(truth, dev(prob), 0.5)
    masks = (prob > 0.5).astype(np.uint8)
    m = R.MeanIoU(2)
    for k in range(B):
        assert imaging.sample_iou_from_counts(counts[k]) == R.sample_iou(truth[k], masks[k])
        assert imaging.sample_iou_from_counts(counts[k]) == imaging.sample_iou(truth[k], masks[k])
        m.update_state(truth[k], masks[k])
    np.testing.assert_array_equal(counts.sum(0).reshape(2, 2), m.cm)
    from unet_b200.keras_api import MeanIoU
    mm = MeanIoU(2)
    mm.add_confusion(counts.sum(0))
    assert float(mm.result()) == pytest.approx(m.result(), abs=1e-12)


def test_adamw_keras_form():
    n = 10007
    w = RNG.standard_normal(n).astype(np.float32)
    m = np.zeros(n, np.float32); v = np.zeros(n, np.float32)
    wd_, md, vd = dev(w), dev(m), dev(v)
    w64, m64, v64 = w.astype(np.float64), m.astype(np.float64), v.astype(np.float64)
    for t in range(1, 4):
        g = RNG.standard_normal(n).astype(np.float32)
        hyper = dev(np.array([2e-3, 1e-4, 0.9, 0.999, 1e-7, t, 1.0, 0.0], np.float32))
        ops.adamw_step(wd_, dev(g), md, vd, hyper)
        w64, m64, v64 = R.adamw_step(w64, g.astype(np.float64), m64, v64, t, *[float(np.float32(h)) for h in (2e-3, 1e-4, 0.9, 0.999, 1e-7)])
    np.testing.assert_allclose(host(wd_), w64, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(host(md), m64, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(host(vd), v64, rtol=1e-5, atol=1e-9)


def test_cast_transpose():
    a = RNG.standard_normal((70, 45)).astype(np.float32)
    d = torch.empty((70, 45), device="cuda", dtype=torch.bfloat16)
    dt_ = torch.empty((45, 70), device="cuda", dtype=torch.bfloat16)
    ops.cast_transpose_bf16(dev(a), d, dt_)
    np.testing.assert_array_equal(host(d), bf16_round(a))
    np.testing.assert_array_equal(host(dt_), bf16_round(a).T)
    f = torch.empty((70, 45), device="cuda")
    ops.cast(d, f)
    np.testing.assert_array_equal(host(f), bf16_round(a))


def test_rejects_bad_arguments():
    from unet_b200._lib import UnetError
    x = torch.zeros((1, 4, 4, 12), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(UnetError):
        ops.bn_act(x, torch.ones(12, device="cuda"), torch.zeros(12, device="cuda"), torch.empty_like(x))   # C % 8 != 0
    A = torch.zeros((16, 12), device="cuda", dtype=torch.bfloat16)
    with pytest.raises(UnetError):
        ops.gemm(A, torch.zeros((16, 12), device="cuda", dtype=torch.bfloat16), torch.zeros((16, 16), device="cuda"),
                 b_trans=True, tensor_core=True)     # lda % 8 != 0


# ------------------------------------------------------------------------------------------------ CLI pre/post-processing
@pytest.mark.parametrize("src,dst", [((540, 960), (256, 256)), ((100, 37), (64, 64)), ((256, 256), (256, 256)), ((31, 500), (48, 16))])
def test_preprocess_matches_cv2(src, dst):
    import cv2
    img = RNG.integers(0, 256, (src[0], src[1], 3), dtype=np.uint8)
    ref = cv2.resize(img.astype(np.float32) / 255.0, (dst[1], dst[0]), interpolation=cv2.INTER_LINEAR)
    out = torch.empty((dst[0], dst[1], 3), device="cuda")
    ops.preprocess_u8(torch.tensor(img, device="cuda"), out)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=2e-6)


@pytest.mark.parametrize("src,dst", [((256, 256), (540, 960)), ((64, 64), (100, 37)), ((32, 32), (32, 32))])
def test_postprocess_matches_cv2(src, dst):
    import cv2
    prob = RNG.random((src[0], src[1], 1)).astype(np.float32)
    prob[8:20, 8:24] = 0.97
    big = cv2.resize(prob, (dst[1], dst[0]), interpolation=cv2.INTER_LINEAR)
    ref = (big > 0.5).astype(np.uint8) * 255
    mask = torch.empty(dst, device="cuda", dtype=torch.uint8)
    pd = torch.zeros((src[0], src[1], 4), device="cuda")
    pd[..., 2] = torch.tensor(prob[..., 0], device="cuda")
    ops.postprocess_mask(pd[..., 2], mask, 0.5)
    got = mask.cpu().numpy()
    assert set(np.unique(got).tolist()) <= {0, 255}
    borderline = np.abs(big - 0.5) < 1e-5            # pixels whose value sits on the threshold may round either way
    assert np.all((got == ref) | borderline) and (got == ref).mean() > 0.9999


# ------------------------------------------------------------------------------------------------ fused conv_block (inference)
@pytest.mark.parametrize("cfg", [(2, 16, 32, 64, 64), (1, 8, 16, 64, 64), (3, 21, 45, 128, 64), (2, 40, 24, 64, 128),
                                 (1, 64, 64, 256, 128), (2, 9, 7, 8, 16), (1, 33, 70, 72, 104)])
def test_sepconv_fused(cfg):
    n, h, w, cin, cout = cfg
    x = RNG.standard_normal((n, h, w, cin)).astype(np.float32)
    wd = RNG.standard_normal((3, 3, cin)).astype(np.float32) / 3
    wp = (RNG.standard_normal((cin, cout)) / np.sqrt(cin)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, cout).astype(np.float32); sh = RNG.standard_normal(cout).astype(np.float32) * 0.3
    d = bf16_round(R.dwconv3x3(bf16_round(x), wd.astype(np.float64)))       # the A operand is bf16, as in the unfused path
    ref = np.maximum((d.reshape(-1, cin) @ bf16_round(wp)) * sc + sh, 0).reshape(n, h, w, cout)
    buf = torch.zeros((n, h, w, cout + 64), device="cuda", dtype=torch.bfloat16)
    xin = torch.zeros((n, h, w, cin + 8), device="cuda", dtype=torch.bfloat16)
    xin[..., 8:] = dev(x, torch.bfloat16)
    ops.sepconv_fused(xin[..., 8:], dev(wd.reshape(9, cin)), dev(wp.T.copy(), torch.bfloat16), buf[..., 64:], scale=dev(sc), shift=dev(sh))
    got = host(buf)
    np.testing.assert_allclose(got[..., 64:], ref, rtol=1.0 / 64, atol=2e-2)
    assert np.all(got[..., :64] == 0)


@pytest.mark.parametrize("cfg", [(2, 16, 32, 64, 64), (1, 24, 20, 128, 64), (2, 10, 38, 64, 128), (1, 64, 64, 128, 128)])
def test_sepconv_fused_with_pool(cfg):
    """MaxPooling2D((2,2)) from the staged tile of the fused conv_block kernel == max over the stored activation, bit-exact;
    ragged patches, channel-sliced destinations"""
    n, h, w, cin, cout = cfg
    x = dev(RNG.standard_normal((n, h, w, cin)), torch.bfloat16)
    wd = dev(RNG.standard_normal((9, cin)) / 3)
    wpt = dev(RNG.standard_normal((cout, cin)) / np.sqrt(cin), torch.bfloat16)
    sh = dev(RNG.standard_normal(cout) * 0.3)
    ybuf = torch.zeros((n, h, w, 2 * cout), device="cuda", dtype=torch.bfloat16)
    y_ref = torch.empty((n, h, w, cout), device="cuda", dtype=torch.bfloat16)
    pbuf = torch.full((n, h // 2, w // 2, cout + 8), 5.0, device="cuda", dtype=torch.bfloat16)
    ops.sepconv_fused(x, wd, wpt, y_ref, shift=sh)
    ops.sepconv_fused(x, wd, wpt, ybuf[..., cout:], shift=sh, pooled=pbuf[..., :cout])
    assert torch.equal(ybuf[..., cout:], y_ref) and bool((ybuf[..., :cout] == 0).all())
    want = y_ref.float().view(n, h // 2, 2, w // 2, 2, cout).amax(dim=(2, 4))
    assert torch.equal(pbuf[..., :cout].float(), want) and bool((pbuf[..., cout:] == 5.0).all())


@pytest.mark.parametrize("classes", [1, 8])
def test_sepconv_fused_with_head(classes):
    n, h, w, cin, cout = 2, 19, 37, 64, 64
    x = RNG.standard_normal((n, h, w, cin)).astype(np.float32)
    wd = RNG.standard_normal((3, 3, cin)).astype(np.float32) / 3
    wp = (RNG.standard_normal((cin, cout)) / np.sqrt(cin)).astype(np.float32)
    sc = RNG.uniform(0.5, 1.5, cout).astype(np.float32); sh = RNG.standard_normal(cout).astype(np.float32) * 0.3
    hw = (RNG.standard_normal((cout, classes)) / 4).astype(np.float32); hb = RNG.standard_normal(classes).astype(np.float32) * 0.1
    d = bf16_round(R.dwconv3x3(bf16_round(x), wd.astype(np.float64)))
    y = np.maximum((d.reshape(-1, cin) @ bf16_round(wp)) * sc + sh, 0)     # the head reads the fp32 activations in registers
    logits = y @ hw.astype(np.float64) + hb
    ref = (R.sigmoid(logits) if classes == 1 else R.softmax(logits)).reshape(n, h, w, classes)
    probs = torch.full((n, h, w, classes), float("nan"), device="cuda")
    ops.sepconv_fused(dev(x, torch.bfloat16), dev(wd.reshape(9, cin)), dev(wp.T.copy(), torch.bfloat16), None, scale=dev(sc), shift=dev(sh),
                      head_w=dev(hw), head_b=dev(hb), head_out=probs)
    np.testing.assert_allclose(host(probs), ref, rtol=0, atol=5e-3)
