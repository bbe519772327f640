"""Generate tests/golden/rdsthead_*.npz from the REAL reference RDSTSR with rdst_dim_modify_mode = 'head'
(rdst_variations.py:288-304; build container only).      python oracle/gen_golden_headmode.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

from utils.param_loader import ParametersLoader            # noqa: E402  (reference)
from networks.rdst_variations import make_RDSTSR           # noqa: E402  (reference)
from synth_weights import fill_state_dict, synth_input     # noqa: E402

INI = "/root/reference/config_files/RDST_E1_OASIS_example_SRx4.ini"
OUT = os.path.join(HERE, "..", "tests", "golden")
CASES = [("rdsthead_2blk_x4_16x24_b2", 2, 4, (2, 1, 16, 24), 31, 32), ("rdsthead_3blk_x2_24x24", 3, 2, (1, 1, 24, 24), 33, 34)]


def main():
    torch.set_num_threads(8)
    for name, blocks, scale, shape, wseed, xseed in CASES:
        p = ParametersLoader(INI)
        for k in ("rdst_dense_layer_depths", "rdst_num_heads", "rdst_window_size", "rdst_rdb_depths"):
            setattr(p, k, list(getattr(p, k))[:blocks])
        p.sr_scale = float(scale)
        p.rdst_dim_modify_mode = "head"
        torch.manual_seed(0)
        m = make_RDSTSR(p).eval()
        m.load_state_dict(fill_state_dict(m.state_dict(), wseed, True), strict=True)
        x = synth_input(shape, xseed)
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy(), shape=np.array(shape),
                            **{"meta_" + k: np.array(v) for k, v in dict(blocks=blocks, scale=scale, wseed=wseed, xseed=xseed).items()})
        with open(os.path.join(OUT, name + "_manifest.txt"), "w") as f:
            for k, v in m.state_dict().items():
                f.write(f"{k}\t{tuple(v.shape)}\t{str(v.dtype).replace('torch.', '')}\n")
        print(f"{name}: out {tuple(y.shape)} min {y.min():.5f} max {y.max():.5f} keys {len(m.state_dict())}")


if __name__ == "__main__":
    main()
