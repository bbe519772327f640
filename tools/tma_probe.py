"""TMA probe: 4x4-token boxes of a shifted 8x8 window, SWIZZLE_128B image in shared memory, store round trip."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdst_b200 import _lib  # noqa: E402


def expect_image(x, C, shift, b, hs0, ws0):
    """[panel][row][128 B] image; row = q*16 + (iy&3)*4 + (ix&3), q = (iy>>2)*2 + (ix>>2); chunk c at c ^ (row & 7)."""
    B, H, W, ld = x.shape
    np_ = (C + 63) // 64
    img = torch.zeros(np_, 64, 8, 8, dtype=torch.bfloat16)
    for iy in range(8):
        for ix in range(8):
            r = ((iy >> 2) * 2 + (ix >> 2)) * 16 + (iy & 3) * 4 + (ix & 3)
            h, w = (hs0 + iy + shift) % H, (ws0 + ix + shift) % W
            v = torch.zeros(np_ * 64, dtype=torch.bfloat16)
            v[:C] = x[b, h, w, :C].cpu()
            for p in range(np_):
                for c in range(8):
                    img[p, r, c ^ (r & 7)] = v[p * 64 + c * 8: p * 64 + c * 8 + 8]
    return img


def main():
    ok = True
    for (C, ld, shift, hs0, ws0) in [(64, 64, 0, 8, 16), (96, 160, 4, 32, 24), (128, 128, 4, 16, 24), (96, 96, 0, 0, 0)]:
        B, H, W = 3, 40, 32
        x = torch.randn(B, H, W, ld, device="cuda").to(torch.bfloat16)
        y = torch.full((B, H, W, 128), 7.0, device="cuda", dtype=torch.bfloat16)
        np_ = (C + 63) // 64
        dump = torch.zeros(np_ * 8192 // 2, device="cuda", dtype=torch.bfloat16)
        _lib.call("rdst_tma_selftest", _lib.ptr(x), ld, _lib.ptr(y), 128, B, H, W, C, shift, 1, hs0, ws0, _lib.ptr(dump),
                  _lib.stream_ptr())
        torch.cuda.synchronize()
        got = dump.cpu().view(np_, 64, 8, 8)
        exp = expect_image(x, C, shift, 1, hs0, ws0)
        e1 = bool((got == exp).all())
        yy = torch.full((B, H, W, 128), 7.0, dtype=torch.bfloat16)
        for iy in range(8):
            for ix in range(8):
                h, w = (hs0 + iy + shift) % H, (ws0 + ix + shift) % W
                yy[1, h, w, :C] = x[1, h, w, :C].cpu()
        e2 = bool((y.cpu() == yy).all())
        print(f"C={C} ld={ld} shift={shift} window@({hs0},{ws0}): image {'OK' if e1 else 'MISMATCH'}, store {'OK' if e2 else 'MISMATCH'}")
        ok = ok and e1 and e2
    print("ALL OK" if ok else "FAILED")


if __name__ == "__main__":
    main()
