import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200 import ops
n, h, w, c = 2, 6, 10, 16
dt = torch.bfloat16
z = torch.randn((n, h, w, c), device="cuda").to(dt)
sc = torch.rand(c, device="cuda") + 0.5; sh = torch.randn(c, device="cuda")
dpool = torch.randn((n, h // 2, w // 2, c), device="cuda").to(dt)
cat = torch.zeros((n, h, w, 2 * c), device="cuda", dtype=dt)
out = torch.empty_like(z)
print("plain", flush=True)
ops.maxpool2x2_bwd(z, sc, sh, dpool, cat[..., c:], out)
torch.cuda.synchronize()
print("drop0", flush=True)
from unet_b200._lib import Dropout
d0 = Dropout(0.0, 23, 2 * c, c, None)
ops.maxpool2x2_bwd(z, sc, sh, dpool, cat[..., c:], out, skip_drop=d0)
torch.cuda.synchronize()
print("drop0 ok", flush=True)
ops.maxpool2x2_bwd(z, sc, sh, dpool, None, out, skip_drop=ops.make_dropout(0.25, 23, ctot=2 * c, c0=c))
torch.cuda.synchronize()
print("drop noskip ok", flush=True)
d = ops.make_dropout(0.25, 23, ctot=2 * c, c0=c)
print(d, d.rate, d.seed, d.ctot, d.c0, d.seed_dev, flush=True)
ops.maxpool2x2_bwd(z, sc, sh, dpool, cat[..., c:], out, skip_drop=d)
torch.cuda.synchronize()
print("ok", flush=True)
g = torch.empty((n * (h // 2) * (w // 2), 4 * c), device="cuda", dtype=dt)
db = torch.zeros(c, device="cuda")
ops.convt_bwd_gather(cat[..., :c], g, db, drop=ops.make_dropout(0.25, 23, ctot=2 * c, c0=0))
torch.cuda.synchronize()
print("ok2", flush=True)
