"""TMEM -> register (tcgen05.ld) and register -> TMEM (tcgen05.st) bandwidth of one SM vs number of warps."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rdst_b200 import _lib
out = torch.zeros(3, dtype=torch.int64, device="cuda")
for store in (0, 1):
    for nw in (1, 2, 4, 8, 16):
        for reps in (1, 16):
            best = None
            for _ in range(3):
                _lib.call("rdst_tmem_bw_bench", nw, reps, store, _lib.ptr(out), _lib.stream_ptr())
                torch.cuda.synchronize()
                t = out.cpu().tolist()
                best = t if best is None or t[0] < best[0] else best
            print(f"{'st' if store else 'ld'} warps={nw:2d} rounds={reps:2d}: {best[0]:6d} cycles, {best[1] / best[0]:7.1f} B/cycle/SM, "
                  f"{best[0] / reps:7.1f} cycles per 16 KB round per warp")
