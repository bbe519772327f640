"""world_size-2 gloo test (CPU) of the data-parallel gradient exchange (rdst_b200/ddp.py): buckets follow the links of the
network, every bucket's all-reduce is launched as soon as its last gradient has been accumulated (reverse link order =
overlap with the rest of backward), `p.grad` are views of the flat bucket buffers, and the averaged gradients equal the
mean of the per-rank gradients computed in one process."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Link(torch.nn.Module):
    def __init__(self, d):
        super().__init__()
        self.fc = torch.nn.Linear(d, d)
        self.norm = torch.nn.LayerNorm(d)

    def forward(self, x):
        return x + self.fc(self.norm(x))


class _Net(torch.nn.Module):
    """Same top-level names as RDSTSR (head / patch_embed / body.i / norm / tail) so rdst_link_of() applies."""

    def __init__(self, d=8, n=3):
        super().__init__()
        self.head = torch.nn.Linear(d, d)
        self.patch_embed = torch.nn.LayerNorm(d)
        self.body = torch.nn.ModuleList([_Link(d) for _ in range(n)])
        self.norm = torch.nn.LayerNorm(d)
        self.tail = torch.nn.Linear(d, 1)
        self.frozen = torch.nn.Parameter(torch.ones(1), requires_grad=False)

    def forward(self, x):
        x = self.patch_embed(self.head(x))
        for b in self.body:
            x = b(x)
        return self.tail(self.norm(x))


def _data(rank):
    g = torch.Generator().manual_seed(10 + rank)
    return torch.randn(16, 8, generator=g), torch.randn(16, 1, generator=g)


def _worker(rank, world, port, ret):
    import sys
    if helpers.ROOT not in sys.path:
        sys.path.insert(0, helpers.ROOT)
    from rdst_b200 import ddp
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    torch.manual_seed(0)
    net = _Net()
    red = ddp.BucketedAllReduce(net)
    x, y = _data(rank)
    ok = True
    for step in range(2):                                   # second step: buffers re-zeroed, views still attached
        red.begin_step()
        torch.nn.functional.l1_loss(net(x), y).backward()
        order = list(red.launch_order)
        red.finish()
        ok &= order == ["tail", "body.2", "body.1", "body.0", "head"]
        ok &= all(p.grad.data_ptr() == v.data_ptr() for b in red.buckets for p, v in zip(b["params"], b["views"]))
    # reference: mean over ranks of the single-process gradients
    torch.manual_seed(0)
    ref = _Net()
    acc = [torch.zeros_like(p) for p in ref.parameters() if p.requires_grad]
    for r in range(world):
        ref.zero_grad()
        xr, yr = _data(r)
        torch.nn.functional.l1_loss(ref(xr), yr).backward()
        for a, p in zip(acc, [p for p in ref.parameters() if p.requires_grad]):
            a += p.grad / world
    got = [p.grad for p in net.parameters() if p.requires_grad]
    err = max((g - a).abs().max().item() for g, a in zip(got, acc))
    # a backward pass that leaves a bucket without gradients must be reported, not silently skipped
    red.begin_step()
    net.body[0](x).sum().backward()
    try:
        red.finish()
        raised = False
    except RuntimeError as e:
        raised = "received no gradient" in str(e)
    for w in red._works:
        w.wait()
    red.remove()
    # allow_unused: the bucket of the unused link is reduced at finish() (zeros), the used ones as usual
    red2 = ddp.BucketedAllReduce(net, allow_unused=True)
    red2.begin_step()
    net.body[0](x).sum().backward()
    red2.finish()
    unused_ok = red2.launch_order[0] == "body.0" and set(red2.launch_order) == {"body.0", "body.1", "body.2", "head", "tail"} \
        and float(net.tail.weight.grad.abs().max()) == 0.0
    red2.remove()
    # a training loop that keeps optimizer.zero_grad(set_to_none=True) instead of begin_step(): autograd assigns fresh gradients,
    # the hooks gather them into the flat buffers; two steps, averaged gradients as above both times
    red3 = ddp.BucketedAllReduce(net)
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.0)
    zg_err = 0.0
    for step in range(2):
        opt.zero_grad(set_to_none=True)
        torch.nn.functional.l1_loss(net(x), y).backward()
        red3.finish()
        got3 = [p.grad for p in net.parameters() if p.requires_grad]
        zg_err = max(zg_err, max((g - a).abs().max().item() for g, a in zip(got3, acc)))
    red3.remove()
    if rank == 0:
        ret.update(ok=bool(ok), err=err, raised=raised, frozen_grad=net.frozen.grad is None, unused_ok=bool(unused_ok),
                   keys=[b["key"] for b in red.buckets], zg_err=zg_err)
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_allreduce_two_ranks():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["ok"], "launch order / gradient views"
    assert ret["err"] < 1e-6, ret["err"]
    assert ret["raised"] and ret["frozen_grad"] and ret["unused_ok"]
    assert sorted(ret["keys"]) == ["body.0", "body.1", "body.2", "head", "tail"]
    assert ret["zg_err"] < 1e-6, ret["zg_err"]      # optimizer.zero_grad() in place of begin_step()


def test_rdst_link_of_maps_state_dict_names():
    from rdst_b200 import ddp
    assert ddp.rdst_link_of("body.3.body.1.body.blocks.0.attn.qkv.weight") == "body.3"
    assert ddp.rdst_link_of("body.0.conv.bias") == "body.0"
    assert ddp.rdst_link_of("head.weight") == "head" and ddp.rdst_link_of("patch_embed.norm.bias") == "head"
    for n in ("norm.weight", "conv_after_body.bias", "tail.0.0.weight", "tail.1.bias", "bottleneck.0.weight", "upsample.0.bias"):
        assert ddp.rdst_link_of(n) == "tail"
    assert ddp.rdst_link_of("layers.2.residual_group.blocks.5.mlp.fc1.weight") == "layers.2"      # SwinIR
    assert ddp.rdst_link_of("layers.0.conv.weight") == "layers.0" and ddp.rdst_link_of("conv_first.bias") == "head"


def _rdst_worker(rank, world, port, ret):
    """The real module (2 RDSTBs) through the real autograd chain, kernels replaced by their contract restatements."""
    import sys
    for p in (helpers.ROOT, os.path.join(helpers.ROOT, "oracle"), os.path.join(helpers.ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from abi_emulator import emulated_abi
    from rdst_b200 import autograd, ddp
    from synth_weights import fill_state_dict
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    m = helpers.make_module(2, 2, "fp32")
    m.load_state_dict(fill_state_dict(helpers.skeleton_state_dict(2, 2), 3, True))
    m.train()
    red = ddp.BucketedAllReduce(m)
    g = torch.Generator().manual_seed(20 + rank)
    x, y = torch.rand(1, 1, 8, 16, generator=g), torch.rand(1, 1, 16, 32, generator=g)
    with emulated_abi():
        red.begin_step()
        torch.nn.functional.l1_loss(autograd.forward_with_grad(m._exec, x), y).backward()
        order = list(red.launch_order)
        red.finish()
    gsum = torch.stack([p.grad.double().sum() for p in m.parameters() if p.requires_grad]).sum()
    lo, hi = gsum.clone(), gsum.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if rank == 0:
        ret.update(order=order, same=bool(lo == hi), finite=bool(torch.isfinite(gsum)))
    dist.barrier()
    dist.destroy_process_group()


def test_rdst_module_buckets_become_ready_block_by_block():
    """The claim behind the all-reduce overlap: with the per-RDSTB autograd Functions and interleaved packing, the
    gradient buckets of the real module complete in reverse link order while backward is still running."""
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_rdst_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["order"] == ["tail", "body.1", "body.0", "head"]
    assert ret["same"] and ret["finite"]
