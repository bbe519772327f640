"""The oracle must reproduce the reference module's own outputs (fixtures from oracle/gen_golden.py)."""
import pytest
import torch

import helpers
import rdst_oracle as O


@pytest.mark.parametrize("name", helpers.CASES)
def test_oracle_matches_reference_golden(name):
    c = helpers.load_case(name)
    taps = {}
    y = O.forward(c["sd"], c["x"], c["scale"], taps=taps)
    assert (y - torch.from_numpy(c["g"]["y"])).abs().max().item() < 2e-5
    for k in ("head", "embed", "rdstb0", "feat"):
        if k in c["g"].files:
            assert (taps[k] - torch.from_numpy(c["g"][k])).abs().max().item() < 2e-5, k


def test_sanity_anchor_shapes_and_mask():
    m = O.shift_mask(16, 24)
    assert m.shape == (6, 64, 64) and set(m.unique().tolist()) == {-100.0, 0.0}
    assert (m[0] == 0).all()                      # interior window carries no masking
    idx = O.rel_pos_index()
    assert idx.min() == 0 and idx.max() == 224 and idx[0, 63] == 0 and idx[63, 0] == 224
    with pytest.raises(ValueError):
        O.to_windows(torch.zeros(1, 12, 16, 4))
