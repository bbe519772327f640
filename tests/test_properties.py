"""Property tests (hypothesis) of the host-side index arithmetic: slice sharding, padded channel layout, packed-buffer layouts."""
import torch
from hypothesis import given, settings, strategies as st

from rdst_b200 import autograd, infer, network, packing


@given(n=st.integers(0, 5000), world=st.integers(1, 16))
@settings(max_examples=200, deadline=None)
def test_shard_ranges_partition_the_slice_axis(n, world):
    ranges = [infer.shard_range(n, world, r) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n
    for (b0, e0), (b1, e1) in zip(ranges, ranges[1:]):
        assert e0 == b1 and b0 <= e0
    sizes = [e - b for b, e in ranges]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@given(j=st.integers(0, 3))
def test_channel_positions_are_injective_and_inside_the_padded_width(j):
    c = packing.EMBED + packing.GROWTH * j
    pos = packing.channel_positions(c)
    assert len(set(pos.tolist())) == c and int(pos.max()) < packing.padded_width(c)
    assert pos[:60].tolist() == list(range(60))
    for g in range(j):                                       # growth group g sits at [64 + 32 g, +30)
        assert pos[60 + 30 * g:90 + 30 * g].tolist() == list(range(64 + 32 * g, 94 + 32 * g))


@given(depth=st.integers(1, 6))
@settings(max_examples=6, deadline=None)
def test_link_layout_slots_are_aligned_and_disjoint(depth):
    layer = torch.nn.Module()
    layer.residual_group = network.BasicLayer(60, (24, 24), depth, 6, 2.0, True, None)
    bs = autograd.rstb_layout(layer)
    end = 0
    for off, shape in bs["slots"]:
        assert off % 4 == 0 and off >= end                   # 16-byte aligned, no overlap
        end = off + autograd._numel(shape)
    assert end <= bs["packed_floats"] and bs["n_params"] == 13 * depth == len(autograd.rstb_params(layer))
    assert len(bs["lins"]) == 4 * depth and len(bs["tables"]) == depth
    assert [s["shift"] for s in bs["stl"]] == [0 if i % 2 == 0 else 4 for i in range(depth)]
    used = sorted(i for l in bs["lins"] for i in (l["w"], l["b"], l["g"], l["be"]) if i is not None) + sorted(bs["tables"])
    assert sorted(used) == list(range(bs["n_params"]))       # every parameter of the link is consumed exactly once
