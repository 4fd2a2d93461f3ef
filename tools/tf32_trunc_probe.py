#!/usr/bin/env python3
"""Does tcgen05.mma.kind::tf32 truncate or round the low 13 mantissa bits of its fp32 operands?  B is handed to the tf32x3
GEMM raw (as its own `hi`) with lo = b - trunc13(b) or lo = b - rna_tf32(b); whichever keeps the product at fp32 accuracy
tells which value the hardware actually multiplies."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from unet_b200 import ops

torch.manual_seed(0)
M, K, N = 4096, 256, 128
a = torch.randn(M, K, device="cuda")
b = torch.randn(N, K, device="cuda") / K ** 0.5
ref = a.double() @ b.double().T
c = torch.empty(M, N, device="cuda")
lo = torch.empty_like(b)
ops.split_tf32(b, None, lo)
ops.gemm(a, b, c, b_trans=True, B_lo=lo, tensor_core=True)
print("library split (raw b + tf32_rna(b - trunc b)): max err", float((c.double() - ref).abs().max()))
hi = torch.empty_like(b)
ops.split_tf32(b, hi, lo)
hi = (b.view(torch.int32) + 0x1000 & ~0x1FFF).view(torch.float32)     # rna-rounded hi
trunc = (b.view(torch.int32) & ~0x1FFF).view(torch.float32)
ops.gemm(a, b.clone(), c, b_trans=True, B_lo=(b - trunc).contiguous(), tensor_core=True)
print("raw B, lo = b-trunc: max err", float((c.double() - ref).abs().max()))
ops.gemm(a, b.clone(), c, b_trans=True, B_lo=(b - hi).contiguous(), tensor_core=True)
print("raw B, lo = b-rna  : max err", float((c.double() - ref).abs().max()))
ops.gemm(a, hi, c, b_trans=True, B_lo=torch.zeros_like(b), tensor_core=True)
print("hi only (1 x tf32 on B): max err", float((c.double() - ref).abs().max()))
