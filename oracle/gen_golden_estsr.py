"""Generate tests/golden/estsr_*.npz from the REAL reference ESTSR (residual-in-residual RDSTBs; build container only).

    python oracle/gen_golden_estsr.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

from networks.rdst_variations import ESTSR                  # noqa: E402  (reference)
from synth_weights import fill_state_dict, synth_input     # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
CASES = [("estsr_2x2_x4_16x16_b2", 2, 2, 4, (2, 1, 16, 16), 12, 13), ("estsr_1x3_x2_8x16", 1, 3, 2, (1, 1, 8, 16), 13, 14)]


def main():
    torch.set_num_threads(8)
    for name, n_rr, n_rd, scale, shape, wseed, xseed in CASES:
        torch.manual_seed(0)
        m = ESTSR(img_size=24, patch_size=1, in_chans=1, sr_scale=scale, embed_dim=60, dense_layer_depths=[2] * n_rr,
                  num_heads=[6] * n_rr, window_size=[8] * n_rr, rdb_depths=[3] * n_rr, rrdb_depths=[n_rd] * n_rr,
                  num_rrdb_blocks=n_rr, mlp_ratio=2., pre_norm=True).eval()
        m.load_state_dict(fill_state_dict(m.state_dict(), wseed, True), strict=True)
        x = synth_input(shape, xseed)
        with torch.no_grad():
            y = m(x)
        meta = dict(n_rr=n_rr, n_rd=n_rd, scale=scale, wseed=wseed, xseed=xseed)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy(), shape=np.array(shape),
                            **{"meta_" + k: np.array(v) for k, v in meta.items()})
        with open(os.path.join(OUT, name + "_manifest.txt"), "w") as f:
            for k, v in m.state_dict().items():
                f.write(f"{k}\t{tuple(v.shape)}\t{str(v.dtype).replace('torch.', '')}\n")
        print(f"{name}: out {tuple(y.shape)} min {y.min():.5f} max {y.max():.5f} keys {len(m.state_dict())}")


if __name__ == "__main__":
    main()
