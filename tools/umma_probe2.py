"""GPU probe 2: disable-output-lane masks and A-operand-from-TMEM (writes gpurun_out/umma_probe2.txt)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdst_b200 import _lib

os.makedirs("gpurun_out", exist_ok=True)
out = open("gpurun_out/umma_probe2.txt", "w")
def p(*a):
    s = " ".join(str(x) for x in a); print(s); out.write(s + "\n"); out.flush()
g = torch.Generator().manual_seed(0)
for N, K in [(64, 16), (64, 32), (32, 64), (16, 64)]:
    A = (torch.randn(128, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    B = (torch.randn(2 * N, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.call("rdst_umma_selftest", _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), N, K, 0, 2, _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = torch.cat([A[:64].float() @ B[:N].float().t(), A[64:].float() @ B[N:].float().t()])
    alt = torch.cat([A[:64].float() @ B[N:].float().t(), A[64:].float() @ B[:N].float().t()])
    p(f"lane-mask N={N} K={K}: err(hypothesis)={(D-ref).abs().max().item():.3e} err(inverted)={(D-alt).abs().max().item():.3e}")
for mn in (0, 1):
    for N, K in [(64, 16), (32, 64), (32, 128), (16, 128), (128, 64)]:
        A = (torch.randn(128, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
        B = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
        D = torch.full((128, N), float("nan"), device="cuda")
        _lib.call("rdst_umma_selftest", _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), N, K, mn, 3, _lib.stream_ptr())
        torch.cuda.synchronize()
        ref = A.float() @ B.float().t()
        p(f"A-from-TMEM N={N} K={K} b_mn={mn}: err={(D-ref).abs().max().item():.3e}")
        if not (D - ref).abs().max().item() < 1e-2:
            p("  D[0,:6]", D[0, :6].tolist()); p("  ref[0,:6]", ref[0, :6].tolist())
