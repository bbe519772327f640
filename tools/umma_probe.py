"""GPU probe: run the UMMA self-test kernel over the operand shapes/layouts the fused kernels rely on and print a
diagnostic table (max error; for M=64 the TMEM lane each logical row landed in).  Writes gpurun_out/umma_probe.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rdst_b200 import _lib  # noqa: E402


def run(N, K, mn, m64, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    A = (torch.randn(128, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    B = (torch.randn(N, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    _lib.call("rdst_umma_selftest", _lib.ptr(A), _lib.ptr(B), _lib.ptr(D), N, K, mn, m64, _lib.stream_ptr())
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    return D, ref


def main():
    os.makedirs("gpurun_out", exist_ok=True)
    out = open("gpurun_out/umma_probe.txt", "w")

    def p(*a):
        s = " ".join(str(x) for x in a)
        print(s)
        out.write(s + "\n")
        out.flush()

    p("device", torch.cuda.get_device_name(0), "tcgen05", _lib.load().rdst_has_tcgen05())
    for mn in (0, 1):
        for N, K in [(16, 16), (32, 16), (64, 64), (128, 16), (128, 32), (128, 128), (192, 96), (240, 128), (256, 256),
                     (64, 128), (16, 128), (32, 128), (96, 64)]:
            try:
                D, ref = run(N, K, mn, 0)
                err = (D - ref).abs().max().item()
                p(f"M=128 N={N:3d} K={K:3d} b_mn_major={mn}: max_err={err:.3e} ref_max={ref.abs().max().item():.2f} "
                  f"{'OK' if err < 1e-2 else 'MISMATCH'}")
                if not err < 1e-2:
                    p("   D[0,:8]  ", D[0, :8].tolist())
                    p("   ref[0,:8]", ref[0, :8].tolist())
                    p("   D[1,:8]  ", D[1, :8].tolist())
                    p("   ref[1,:8]", ref[1, :8].tolist())
                    # hypothesis checks
                    if N == 128:
                        p("   transposed?", (D - ref.t()).abs().max().item())
            except Exception as e:  # noqa: BLE001
                p(f"M=128 N={N} K={K} mn={mn}: EXCEPTION {e}")
    # M=64: where do rows land?
    for N, K in [(64, 16), (64, 64), (16, 64)]:
        try:
            D, ref = run(N, K, 0, 1)
            ref64 = ref[:64]
            lanes = []
            for r in range(64):
                d = (D - ref64[r][None]).abs().max(1).values
                d = torch.nan_to_num(d, nan=1e9)
                l = int(d.argmin())
                lanes.append((r, l, float(d[l])))
            bad = [x for x in lanes if x[2] > 1e-2]
            p(f"M=64 N={N} K={K}: row->lane", [(r, l) for r, l, _ in lanes][:64])
            p(f"   unmatched rows: {len(bad)}")
        except Exception as e:  # noqa: BLE001
            p(f"M=64 N={N} K={K}: EXCEPTION {e}")
    out.close()


if __name__ == "__main__":
    main()
