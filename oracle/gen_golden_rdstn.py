"""Generate tests/golden/rdstn_*.npz from the REAL reference RDSTSR_N (global bottleneck 'mlp' / 'conv'; build container only).

    python oracle/gen_golden_rdstn.py
"""
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
CASES = [("rdstn_e1_x4_40x32", 8, 4, (1, 1, 40, 32), 8, 9, "mlp"), ("rdstn_2blk_x2_16x24_b2", 2, 2, (2, 1, 16, 24), 9, 10, "mlp"),
         ("rdstn_conv_3blk_x4_16x16", 3, 4, (2, 1, 16, 16), 10, 11, "conv")]


def main():
    torch.set_num_threads(8)
    for name, blocks, scale, shape, wseed, xseed, mode in CASES:
        p = ParametersLoader(INI)
        for k in ("rdst_dense_layer_depths", "rdst_num_heads", "rdst_window_size", "rdst_rdb_depths"):
            setattr(p, k, list(getattr(p, k))[:blocks])
        p.sr_scale = float(scale)
        p.rdst_global_bottleneck, p.rdst_global_bottleneck_mode = True, mode
        torch.manual_seed(0)
        m = make_RDSTSR(p).eval()
        assert type(m).__name__ == "RDSTSR_N"
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
