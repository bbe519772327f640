import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdst_b200 import _lib
g = torch.Generator().manual_seed(0)
N, K = 64, 16
A = (torch.randn(128, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
B = (torch.randn(2 * N, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
D = torch.full((128, N), float("nan"), device="cuda")
_lib.call("rdst_umma_selftest", _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), N, K, 0, 2, _lib.stream_ptr())
torch.cuda.synchronize()
c0 = A.float() @ B[:N].float().t()
c1 = A.float() @ B[N:].float().t()
cls = []
for r in range(128):
    d = D[r]
    if torch.isnan(d).all(): cls.append('n')
    elif (d - c0[r]).abs().max() < 1e-3: cls.append('0')
    elif (d - c1[r]).abs().max() < 1e-3: cls.append('1')
    elif (d - c0[r] - c1[r]).abs().max() < 1e-3: cls.append('+')
    else:
        # column-wise classification
        m0 = ((d - c0[r]).abs() < 1e-3); m1 = ((d - c1[r]).abs() < 1e-3)
        cls.append('[' + ''.join('0' if a else ('1' if b else '?') for a, b in zip(m0.tolist(), m1.tolist())) + ']')
print(''.join(cls))
print("D[64,:6]", D[64,:6].tolist())
print("c0[64,:6]", c0[64,:6].tolist())
print("c1[64,:6]", c1[64,:6].tolist())
print("D[100,:6]", D[100,:6].tolist())
print("c1[100,:6]", c1[100,:6].tolist())
# does any row of c0/c1 match D[64]?
for name, c in (("c0", c0), ("c1", c1)):
    d = (c - D[64][None]).abs().max(1).values
    print(name, "best row for D[64]:", int(d.argmin()), float(d.min()))
# run again (TMEM now holds previous results) to see if values are stale
D2 = torch.full((128, N), float("nan"), device="cuda")
_lib.call("rdst_umma_selftest", _lib.ptr(A), _lib.ptr(B), _lib.ptr(D2), N, K, 0, 2, _lib.stream_ptr())
torch.cuda.synchronize()
print("second run equal to first on rows 64+:", bool(torch.equal(D[64:], D2[64:])))
# normal run then masked run: do rows 64+ keep the normal run's values?
D3 = torch.full((128, N), float("nan"), device="cuda")
_lib.call("rdst_umma_selftest", _lib.ptr(A), _lib.ptr(B), _lib.ptr(D3), N, K, 0, 0, _lib.stream_ptr())
_lib.call("rdst_umma_selftest", _lib.ptr(A), _lib.ptr(B), _lib.ptr(D2), N, K, 0, 2, _lib.stream_ptr())
torch.cuda.synchronize()
print("rows 64+ after normal-then-masked == normal run (stale)?", float((D2[64:] - D3[64:]).abs().max()))
