"""Three launches of the training GEMM kernels at the cfg4 shape for an ncu capture:
   ncu --set full --clock-control none --import-source on -k regex:gemm_t -o gpurun_out/train_gemm python tools/ncu_train_gemm.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdst_b200 import _lib as L
T, cp, c = 18432, 128, 120
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(T, cp, device="cuda", generator=g); x[:, c:] = 0
w = torch.randn(3 * c, cp, device="cuda", generator=g) * 0.05
b = torch.randn(3 * c, device="cuda", generator=g)
y = torch.empty(T, 3 * c, device="cuda")
dx = torch.empty(T, cp, device="cuda")
dw = torch.zeros(3 * c, cp, device="cuda"); db = torch.zeros(3 * c, device="cuda")
st = L.stream_ptr()
for _ in range(2):
    # qkv forward (LayerNorm prologue), qkv data gradient (MN-major weight), qkv weight gradient (LayerNorm-hat operand)
    L.call("rdst_gemm_tc", L.ptr(x), cp, L.ptr(w), cp, 0, L.ptr(b), None, 0, None, 0, L.ptr(y), 3 * c, T, cp, 3 * c, 1, c, 1.0, 0, 0, 0, 0, 0, 0, st)
    L.call("rdst_gemm_tc", L.ptr(y), 3 * c, L.ptr(w), cp, 1, None, None, 0, None, 0, L.ptr(dx), cp, T, 3 * c, cp, 0, 0, 1.0, 0, 0, 0, 0, 0, 0, st)
    L.call("rdst_gemm_tn_tc", L.ptr(y), 3 * c, L.ptr(x), cp, L.ptr(dw), L.ptr(db), T, 3 * c, cp, 0, 0, 0, 0, 0, 1, c, st)
torch.cuda.synchronize()
