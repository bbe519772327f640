"""Training path (forward with autograd graph + hand-written backward kernels)."""


def forward_with_grad(executor, x):
    raise NotImplementedError(
        "rdst_b200: the backward kernels are not built yet; run inference under torch.no_grad() "
        "(there is deliberately no PyTorch-autograd fallback).")
