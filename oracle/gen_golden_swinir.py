"""Generate tests/golden/swinir_*.npz from the REAL reference SwinIR (build container only; needs /root/reference).

    python oracle/gen_golden_swinir.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

from utils.param_loader import ParametersLoader                      # noqa: E402  (reference)
from networks.swin_transformer_sr import SwinIR, swinir_make_model   # noqa: E402  (reference)
from synth_weights import fill_state_dict, synth_input               # noqa: E402

INI = "/root/reference/config_files/RDST_E1_OASIS_example_SRx4.ini"
OUT = os.path.join(HERE, "..", "tests", "golden")

CASES = [
    # name,                 factory / ctor,                                  input shape,     wseed, xseed
    ("swinir_ini_x4_40x32", None,                                            (1, 1, 40, 32),  5,     6),   # the ini's SwinIR-lite
    ("swinir_x2_16x24_b2",  dict(img_size=24, depths=[2, 2], upscale=2),     (2, 1, 16, 24),  6,     7),   # shifted blocks
    ("swinir_x3_8x8",       dict(img_size=24, depths=[2], upscale=3),        (3, 1, 8, 8),    7,     8),   # one window, shift
]


def build(spec):
    torch.manual_seed(0)
    if spec is None:
        return swinir_make_model(ParametersLoader(INI)).eval()
    return SwinIR(img_size=spec["img_size"], patch_size=1, in_chans=1, embed_dim=60, depths=spec["depths"],
                  num_heads=[6] * len(spec["depths"]), window_size=8, mlp_ratio=2., upscale=spec["upscale"], img_range=1.,
                  upsampler="pixelshuffledirect", resi_connection="1conv").eval()


def main():
    torch.set_num_threads(8)
    for name, spec, shape, wseed, xseed in CASES:
        m = build(spec)
        m.load_state_dict(fill_state_dict(m.state_dict(), wseed, True), strict=True)
        x = synth_input(shape, xseed)
        with torch.no_grad():
            y = m(x)
        meta = dict(img_size=8 if spec is None else spec["img_size"], upscale=m.upscale, wseed=wseed, xseed=xseed)
        depths = [6, 6, 6, 6] if spec is None else spec["depths"]
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy(), shape=np.array(shape), depths=np.array(depths),
                            **{"meta_" + k: np.array(v) for k, v in meta.items()})
        with open(os.path.join(OUT, name + "_manifest.txt"), "w") as f:
            for k, v in m.state_dict().items():
                f.write(f"{k}\t{tuple(v.shape)}\t{str(v.dtype).replace('torch.', '')}\n")
        print(f"{name}: out {tuple(y.shape)} min {y.min():.5f} max {y.max():.5f} mean {y.mean():.6f} keys {len(m.state_dict())}")


if __name__ == "__main__":
    main()
