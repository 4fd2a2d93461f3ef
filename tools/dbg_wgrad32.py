import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_b200 import ops
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
P, M, N = 32, 128, 64
a = torch.zeros(P, M, device="cuda"); b = torch.zeros(P, N, device="cuda")
for p in range(P):
    a[p, p] = 1.0
    a[p, 64 + p] = 2.0
b[:, :] = torch.arange(1, N + 1, device="cuda").float()[None, :] + 100 * torch.arange(P, device="cuda").float()[:, None]
c = torch.zeros(M, N, device="cuda")
ops.gemm(a, b, c, a_trans=True, accumulate=True, tf32x3=True, tensor_core=True)
torch.cuda.synchronize()
ref = a.double().T @ b.double()
print("max |c|", float(c.abs().max()), "max err", float((c.double() - ref).abs().max()))
print(c[:4, :8]); print(ref[:4, :8].float())
print(c[64:68, :8]); print(c[32:36, :8])
nz = (c != 0).nonzero()
print("nonzero count", nz.shape[0], nz[:10].tolist())
