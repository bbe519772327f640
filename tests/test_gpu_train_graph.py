"""GraphedTrainStep: the whole training step (forward, L1, backward, optimizer) replayed as one CUDA graph must
reproduce eager training from the same initial weights; warm-up / capture steps must not leak into the weights."""
import copy

import pytest
import torch

import helpers

pytestmark = pytest.mark.gpu


def _opt(m, capturable):
    return torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.99), eps=1e-8,
                            capturable=capturable, fused=True)


@pytest.mark.parametrize("kind,precision", [("rdst", "fp32"), ("rdst", "bf16"), ("swinir", "bf16")])
def test_graph_replay_equals_eager(kind, precision):
    from rdst_b200 import train, ddp
    torch.manual_seed(0)
    build = (lambda: helpers.make_module(1, 4, precision)) if kind == "rdst" else \
        (lambda: helpers.make_swinir(dict(img_size=24, depths=[2], upscale=4), precision))
    m = build().cuda().train()
    sd0 = copy.deepcopy(m.state_dict())
    g = torch.Generator(device="cuda").manual_seed(1)
    batches = [(torch.rand(2, 1, 16, 24, device="cuda", generator=g), torch.rand(2, 1, 64, 96, device="cuda", generator=g))
               for _ in range(3)]
    red = ddp.BucketedAllReduce(m)                        # world size 1: buckets, flat gradients, no collective
    step = train.GraphedTrainStep(m, _opt(m, True), *batches[0], reducer=red)
    for p, q in zip(m.state_dict().values(), sd0.values()):
        assert torch.equal(p, q)                          # capture left the weights untouched
    losses = [float(step(x, y)) for x, y in batches]
    assert red.launch_order[0] == "tail" and red.launch_order[-1] == "head"
    red.remove()
    m2 = build().cuda().train()
    m2.load_state_dict(sd0)
    opt2 = _opt(m2, False)
    losses2 = []
    for x, y in batches:
        opt2.zero_grad(set_to_none=True)
        loss = torch.nn.functional.l1_loss(m2(x), y)
        loss.backward()
        opt2.step()
        losses2.append(float(loss))
    tol = 1e-5 if precision == "fp32" else 2e-4           # fp32 atomics in the weight-gradient reductions reorder sums
    assert max(abs(a - b) for a, b in zip(losses, losses2)) < 10 * tol
    diff = max((p.detach() - q.detach()).abs().max().item() for p, q in zip(m.parameters(), m2.parameters()))
    assert diff < 100 * tol, diff


def test_optimizer_must_be_capturable():
    from rdst_b200 import train
    m = helpers.make_module(1, 4, "fp32").cuda().train()
    x, y = torch.rand(1, 1, 8, 8, device="cuda"), torch.rand(1, 1, 32, 32, device="cuda")
    with pytest.raises(ValueError, match="capturable"):
        train.GraphedTrainStep(m, torch.optim.Adam(m.parameters(), lr=1e-3), x, y)
