"""Gradient parity of the training path: rdst_b200 (fp32 kernels, hand-written backward) vs torch.autograd through the
CPU oracle on identical weights / inputs / L1 loss (the reference's training objective, loss/sr_loss.py RecLoss('L1'))."""
import pytest
import torch

import helpers
import rdst_oracle as O

pytestmark = pytest.mark.gpu


def _oracle_grads(sd, x, target, scale):
    p = {k: (v.clone().double().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    out = O.forward(p, x.double(), scale)
    loss = (out - target.double()).abs().mean()
    names = [k for k, v in p.items() if v.is_floating_point() and "mean." not in k and "attn_mask" not in k]
    grads = torch.autograd.grad(loss, [p[k] for k in names], allow_unused=True)
    return float(loss), dict(zip(names, grads)), out.detach()


@pytest.mark.parametrize("blocks,shape,scale", [(1, (2, 1, 16, 24), 4), (2, (1, 1, 24, 24), 2), (1, (3, 1, 8, 8), 4)])
def test_gradients_match_oracle_autograd(blocks, shape, scale):
    c = helpers.load_case("e2blk_x4_8x8")
    from synth_weights import fill_state_dict
    sd = fill_state_dict(helpers.skeleton_state_dict(blocks, scale), 11, True)
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(4))
    target = torch.rand(shape[0], 1, shape[2] * scale, shape[3] * scale, generator=torch.Generator().manual_seed(5))
    loss_ref, g_ref, out_ref = _oracle_grads(sd, x, target, scale)

    m = helpers.make_module(blocks, scale, "fp32").cuda().train()
    m.load_state_dict(sd, strict=True)
    out = m(x.cuda())
    assert out.requires_grad
    loss = torch.nn.functional.l1_loss(out, target.cuda())
    loss.backward()
    assert abs(float(loss) - loss_ref) < 1e-5
    assert (out.detach().cpu().double() - out_ref).abs().max().item() < 1e-4
    params = dict(m.named_parameters())
    worst = ("", 0.0)
    for k, gr in g_ref.items():
        p = params[k]
        assert p.grad is not None, k
        if gr is None:
            continue
        denom = gr.abs().max().item() + 1e-12
        err = (p.grad.detach().cpu().double() - gr).abs().max().item() / denom
        if err > worst[1]:
            worst = (k, err)
        assert err < 2e-3, (k, err, denom)
    assert m.sub_mean.weight.grad is None and m.add_mean.bias.grad is None
    print("worst relative gradient error:", worst)


def test_training_step_reduces_loss():
    """A few Adam steps (reference optimiser settings, ini :130-135) on one synthetic batch must reduce the L1 loss."""
    torch.manual_seed(0)
    m = helpers.make_module(1, 4, "fp32").cuda().train()
    opt = torch.optim.Adam([p for p in m.parameters() if p.requires_grad], lr=1e-3, betas=(0.9, 0.99), eps=1e-8)
    x = torch.rand(4, 1, 24, 24, device="cuda")
    y = torch.nn.functional.interpolate(x, scale_factor=4, mode="bilinear")
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss = torch.nn.functional.l1_loss(m(x), y)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0] * 0.9, losses
    with torch.no_grad():
        m.eval()
        assert torch.isfinite(m(x)).all()
