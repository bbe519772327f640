"""Generate tests/golden/*.npz from the REAL reference module.  Runs only in the build container
(needs /root/reference); the fixtures it writes are committed and travel to the GPU box.

    python oracle/gen_golden.py

Each case = reference `make_RDSTSR(paras)` (networks/rdst_variations.py:1369) with deterministic weights
from oracle/synth_weights.py, a deterministic input, and the reference output plus a few intermediates
captured with forward hooks (head conv, patch_embed+LN, each RDSTB, conv_after_body).
"""
import copy
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

CASES = [
    # name,            blocks, scale, input shape,      wseed, perturbed, xseed
    ("e1_x4_64x64",      8,     4,  (1, 1, 64, 64),      0,    True,      1),   # BASELINE cfg1 shape
    ("e1_x4_16x24_b2",   8,     4,  (2, 1, 16, 24),      3,    False,     2),   # ragged, on-the-fly mask
    ("e_x4_40x32",       4,     4,  (1, 1, 40, 32),      1,    True,      3),   # RDST-E, OASIS slice shape
    ("e1_x2_24x24",      8,     2,  (1, 1, 24, 24),      2,    True,      4),   # x2 tail, stored-mask path
    ("e2blk_x4_8x8",     2,     4,  (3, 1, 8, 8),        4,    True,      5),   # single window: shift fully masked
]


def build(blocks, scale):
    p = ParametersLoader(INI)
    for name in ("rdst_dense_layer_depths", "rdst_num_heads", "rdst_window_size", "rdst_rdb_depths"):
        setattr(p, name, list(getattr(p, name))[:blocks])
    p.sr_scale = float(scale)
    torch.manual_seed(0)
    return make_RDSTSR(p).eval()


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    for name, blocks, scale, shape, wseed, pert, xseed in CASES:
        m = build(blocks, scale)
        m.load_state_dict(fill_state_dict(m.state_dict(), wseed, pert), strict=True)
        x = synth_input(shape, xseed)
        taps = {}
        hooks = [m.head.register_forward_hook(lambda mod, i, o: taps.__setitem__("head", o.detach().clone())),
                 m.patch_embed.register_forward_hook(lambda mod, i, o: taps.__setitem__("embed", o.detach().clone())),
                 m.conv_after_body.register_forward_hook(lambda mod, i, o: taps.__setitem__("cab", o.detach().clone()))]
        for i, blk in enumerate(m.body):
            hooks.append(blk.register_forward_hook(
                lambda mod, inp, o, i=i: taps.__setitem__(f"rdstb{i}", o.detach().clone())))
        with torch.no_grad():
            y = m(x)
        for h in hooks:
            h.remove()
        keep = {"head": taps["head"], "embed": taps["embed"], "rdstb0": taps["rdstb0"],
                f"rdstb{blocks - 1}": taps[f"rdstb{blocks - 1}"], "cab": taps["cab"]}
        if shape[0] * shape[2] * shape[3] > 1000:      # big case: output only (keeps fixtures small)
            keep = {}
        meta = dict(blocks=blocks, scale=scale, wseed=wseed, perturbed=int(pert), xseed=xseed)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy(), shape=np.array(shape),
                            **{k: v.numpy() for k, v in keep.items()},
                            **{"meta_" + k: np.array(v) for k, v in meta.items()})
        print(f"{name}: out {tuple(y.shape)} min {y.min():.5f} max {y.max():.5f} mean {y.mean():.6f}"
              f"  |rdstb_last| max {taps[f'rdstb{blocks - 1}'].abs().max():.3f}")
    # key/shape manifest of the E1 state_dict: the drop-in's wire format
    m = build(8, 4)
    with open(os.path.join(OUT, "e1_state_dict_manifest.txt"), "w") as f:
        for k, v in m.state_dict().items():
            f.write(f"{k}\t{tuple(v.shape)}\t{str(v.dtype).replace('torch.', '')}\n")
    print("manifest:", len(m.state_dict()), "keys")


if __name__ == "__main__":
    main()
