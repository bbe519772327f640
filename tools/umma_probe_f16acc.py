"""Probe: tcgen05.mma kind::f16 with fp16 accumulators -- where do the N output values of a row land in TMEM?"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdst_b200 import _lib
torch.manual_seed(0)
N, K = 32, 64
A = (torch.randn(128, K) * 0.5).to(torch.float16).cuda()
B = (torch.randn(N, K) * 0.5).to(torch.float16).cuda()
D = torch.zeros(128, N, device="cuda")
_lib.call("rdst_umma_selftest", _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), N, K, 0, 4, _lib.stream_ptr())
torch.cuda.synchronize()
raw = D.view(torch.int32).cpu()
lo = (raw & 0xFFFF).to(torch.int16).view(torch.float16).float()
hi = ((raw >> 16) & 0xFFFF).to(torch.int16).view(torch.float16).float()
ref = A.float().cpu() @ B.float().cpu().t()
print("ref row0[:8]   ", [round(v, 3) for v in ref[0, :8].tolist()])
print("col lo row0[:8]", [round(v, 3) for v in lo[0, :8].tolist()])
print("col hi row0[:8]", [round(v, 3) for v in hi[0, :8].tolist()])
packed = torch.stack([lo[:, :N // 2], hi[:, :N // 2]], dim=-1).reshape(128, N)
print("packed-pairs hypothesis (col c = elements 2c, 2c+1): max err", (packed - ref).abs().max().item())
print("one-per-column hypothesis (low half): max err", (lo - ref).abs().max().item())
