"""GPU parity of the whole drop-in module against the reference goldens and the oracle."""
import pytest
import torch

import helpers
import rdst_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", helpers.CASES)
def test_fp32_matches_reference_golden(name):
    """north_star: fp32 mode within 1e-4 max abs of the reference output on [0,1] images."""
    c = helpers.load_case(name)
    m = helpers.make_module(c["blocks"], c["scale"], "fp32").cuda().eval()
    m.load_state_dict(c["sd"], strict=True)
    with torch.no_grad():
        y = m(c["x"].cuda())
    ref = torch.from_numpy(c["g"]["y"])
    assert y.shape == ref.shape and y.dtype == torch.float32
    assert (y.cpu() - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("name", ["e1_x4_64x64", "e_x4_40x32", "e1_x4_16x24_b2"])
def test_bf16_matches_reference_golden(name):
    """north_star: bf16 mode within 1e-2 max abs and 0.01 dB PSNR of the reference fp32 output."""
    c = helpers.load_case(name)
    m = helpers.make_module(c["blocks"], c["scale"], "bf16").cuda().eval()
    m.load_state_dict(c["sd"], strict=True)
    with torch.no_grad():
        y = m(c["x"].cuda()).cpu()
    ref = torch.from_numpy(c["g"]["y"])
    assert (y - ref).abs().max().item() < 1e-2
    target = helpers.realistic_target(ref)
    assert abs(O.psnr(y, target) - O.psnr(ref, target)) < 0.01


def test_fp32_oracle_at_oasis_batch():
    """Oracle comparison at the OASIS slice shape with a small batch (sizes the oracle finishes in seconds)."""
    c = helpers.load_case("e1_x4_64x64")
    x = torch.rand(4, 1, 40, 32, generator=torch.Generator().manual_seed(5))
    ref = O.forward(c["sd"], x, 4)
    m = helpers.make_module(8, 4, "fp32").cuda().eval()
    m.load_state_dict(c["sd"])
    with torch.no_grad():
        y = m(x.cuda()).cpu()
    assert (y - ref).abs().max().item() < 1e-4


def test_batch_independence_full_volume():
    """Size-independent property at the BASELINE size: slices never interact, so a 176-slice batch must equal
    the same slices run in chunks (bit-exact: same kernels, same per-slice arithmetic)."""
    c = helpers.load_case("e1_x4_64x64")
    m = helpers.make_module(8, 4, "fp32").cuda().eval()
    m.load_state_dict(c["sd"])
    x = torch.rand(176, 1, 40, 32, generator=torch.Generator().manual_seed(9)).cuda()
    with torch.no_grad():
        y = m(x)
        parts = torch.cat([m(x[i:i + 44]).clone() for i in range(0, 176, 44)])
    assert torch.equal(y, parts)
    assert torch.isfinite(y).all()


def test_large_input_matches_oracle_128():
    """cfg5 shape (LR 128x128) in fp32 against the oracle; bf16 against the same output within the bf16 bar."""
    c = helpers.load_case("e2blk_x4_8x8")
    x = torch.rand(1, 1, 128, 128, generator=torch.Generator().manual_seed(21))
    ref = O.forward(c["sd"], x, 4)
    for prec, tol in (("fp32", 1e-4), ("bf16", 1e-2)):
        m = helpers.make_module(c["blocks"], 4, prec).cuda().eval()
        m.load_state_dict(c["sd"])
        with torch.no_grad():
            y = m(x.cuda()).cpu()
        assert (y - ref).abs().max().item() < tol, prec


def test_forward_is_cuda_graph_capturable():
    """The C ABI promises stream-ordered, allocation-free kernels: a captured graph must replay to the same result."""
    c = helpers.load_case("e2blk_x4_8x8")
    m = helpers.make_module(c["blocks"], 4, "bf16").cuda().eval()
    m.load_state_dict(c["sd"])
    x = torch.rand(4, 1, 16, 16, device="cuda")
    with torch.no_grad():
        eager = m(x).clone()          # also builds the packed-weight cache and the workspace outside the capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            m(x)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = m(x)
        x.copy_(torch.rand(4, 1, 16, 16, device="cuda"))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, m(x))
    assert torch.isfinite(eager).all()


def test_graphed_inference_driver_matches_eager():
    from rdst_b200.infer import GraphedRDST, super_resolve_slices
    c = helpers.load_case("e2blk_x4_8x8")
    m = helpers.make_module(c["blocks"], 4, "bf16").cuda().eval()
    m.load_state_dict(c["sd"])
    gm = GraphedRDST(m)
    for shape in ((1, 1, 40, 32), (3, 1, 16, 16), (1, 1, 40, 32)):
        x = torch.rand(*shape, device="cuda")
        with torch.no_grad():
            assert torch.equal(gm(x), m(x))
    vol = torch.rand(10, 1, 40, 32).pin_memory()
    out = super_resolve_slices(m, vol, batch_size=4)
    with torch.no_grad():
        ref = m(vol.cuda()).cpu()
    assert out.shape == (10, 1, 160, 128) and torch.equal(out, ref)
    # graph replay per batch shape (the sharded driver's mode for small per-rank shares): same bits, host and device inputs
    from rdst_b200.infer import super_resolve_volume_sharded
    for src in (vol, vol.cuda()):
        for _ in range(2):
            og = super_resolve_slices(m, src, batch_size=4, use_graph=True)
            assert og.is_cuda == src.is_cuda and torch.equal(og.cpu(), ref)
    b0, b1, part = super_resolve_volume_sharded(m, vol.cuda(), rank=1, world_size=3)
    assert (b0, b1) == (4, 7) and torch.equal(part.cpu(), ref[4:7])


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_out_argument_device_and_pinned_host(precision):
    """forward(x, out=...) writes the HR image into the given tensor -- device memory or pinned host memory -- with the
    same bits as the default path; bad `out` tensors are rejected."""
    c = helpers.load_case("e2blk_x4_8x8")
    m = helpers.make_module(c["blocks"], c["scale"], precision).cuda().eval()
    m.load_state_dict(c["sd"])
    x = torch.rand(5, 1, 16, 24, generator=torch.Generator().manual_seed(2)).cuda()
    with torch.no_grad():
        ref = m(x)
        dev_out = torch.zeros_like(ref)
        assert m(x, out=dev_out) is dev_out and torch.equal(dev_out, ref)
        host_out = torch.zeros(ref.shape).pin_memory()
        m(x, out=host_out)
        torch.cuda.synchronize()
        assert torch.equal(host_out, ref.cpu())
        with pytest.raises(ValueError, match="out="):
            m(x, out=torch.zeros(ref.shape))                      # pageable host memory
        with pytest.raises(ValueError, match="out="):
            m(x, out=torch.zeros(5, 1, 8, 8, device="cuda"))
    from rdst_b200 import infer
    vol = x.cpu().pin_memory()
    y = infer.super_resolve_slices(m, vol, batch_size=2)
    assert y.is_pinned() and torch.equal(y, ref.cpu())


@pytest.mark.parametrize("family", ["rdstn", "estsr", "rdst_many_shapes"])
def test_graph_replay_survives_workspace_turnover(family):
    """graphs captured by GraphedRDST address executor workspaces by raw pointer.  Replaying an earlier graph
    after other shapes were run (RDSTSR_N / ESTSR auxiliary buffers; more shapes than the executor's workspace cache
    holds) must still give the eager result."""
    from rdst_b200.infer import GraphedRDST
    if family == "rdstn":
        c = helpers.load_rdstn_case("rdstn_2blk_x2_16x24_b2")
        m = helpers.make_rdstn(c, "bf16")
    elif family == "estsr":
        c = helpers.load_estsr_case("estsr_2x2_x4_16x16_b2")
        m = helpers.make_estsr(c, "bf16")
    else:
        c = helpers.load_case("e2blk_x4_8x8")
        m = helpers.make_module(c["blocks"], c["scale"], "bf16")
    m = m.cuda().eval()
    m.load_state_dict(c["sd"])
    gm = GraphedRDST(m)
    shapes = [(3, 1, 16, 24), (1, 1, 16, 24)] if family != "rdst_many_shapes" else [(b, 1, 8, 16) for b in range(1, 12)]
    xs = [torch.rand(*s, device="cuda") for s in shapes]
    with torch.no_grad():
        first = [gm(x).clone() for x in xs]
        junk = [torch.randn(1 << 20, device="cuda") for _ in range(8)]      # churn the caching allocator
        for x, y in zip(xs, first):                                          # replay every captured graph again
            assert torch.equal(gm(x), y)
            assert torch.equal(m(x), y)
    del junk
