"""Deterministic synthetic weights for parity tests.  TEST INFRASTRUCTURE ONLY.

The reference ships no checkpoints we can reach (README Dropbox links), so parity tests fill a
reference-format state_dict with values that depend only on (key name, seed).  The same function is used by
oracle/gen_golden.py (reference side, build container) and by tests/ (oracle + CUDA side), so weights never
need to be committed.  "perturbed" mode deliberately moves biases / LayerNorm affine / the relative-position
table far from their init values (0 / 1 / N(0,.02)) so that init symmetries cannot hide indexing bugs.
"""
import zlib

import numpy as np
import torch

_KEEP = ("relative_position_index", "attn_mask", "sub_mean.", "add_mean.")


def _rng(name, seed):
    return np.random.default_rng((zlib.crc32(name.encode()) + 7919 * seed) & 0xFFFFFFFF)


def fill_state_dict(sd, seed=0, perturbed=True):
    """Return a new dict with the same keys/shapes/dtypes as `sd`, floats replaced deterministically."""
    out = {}
    for k, v in sd.items():
        if any(s in k for s in _KEEP) or not v.is_floating_point():
            out[k] = v.clone()
            continue
        r = _rng(k, seed)
        n = r.standard_normal(tuple(v.shape)).astype(np.float32)
        leaf = k.rsplit(".", 1)[-1]
        is_norm = ("norm" in k) or (".tail.0." in k and v.ndim == 1 and "body." in k)
        if "relative_position_bias_table" in k:
            t = n * (0.5 if perturbed else 0.02)
        elif is_norm and leaf == "weight":
            t = 1.0 + (0.1 * n if perturbed else 0.0 * n)
        elif is_norm and leaf == "bias":
            t = (0.05 if perturbed else 0.0) * n
        elif leaf == "bias":
            t = (0.05 if perturbed else 0.0) * n
        elif v.ndim == 4:                       # conv weight: kaiming-uniform-like magnitude
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            t = n * (1.0 / np.sqrt(3.0 * fan_in))
        else:                                   # linear weight
            t = n * (0.04 if perturbed else 0.02)
        out[k] = torch.from_numpy(np.ascontiguousarray(t)).to(v.dtype)
    return out


def synth_input(shape, seed=1):
    """Uniform [0,1) LR input, fp32, deterministic in (shape, seed)."""
    r = np.random.default_rng(1000003 * seed + 17)
    return torch.from_numpy(r.random(tuple(shape), dtype=np.float32))
