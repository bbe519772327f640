"""Data-parallel gradient exchange for RDST training (BASELINE cfg4): one bucket per link of the network, all-reduced
over NCCL (NVLink 5 / NVSwitch) as soon as the bucket's last gradient has been accumulated, i.e. while the backward
kernels of the earlier RDSTBs are still running.

The training path (rdst_b200/autograd.py) is a chain of autograd.Functions (head, RDSTB 0..n-1, tail) with the weight
packing of each link interleaved, so the engine finishes the parameter gradients link by link in reverse order.
`BucketedAllReduce` keeps every bucket's gradients in ONE flat fp32 buffer, counts accumulations with
post-accumulate-grad hooks and, when a bucket's last gradient has landed, gathers the bucket with one multi-tensor copy
(`torch._foreach_copy_`), re-points `p.grad` at the views of the flat buffer (what the optimizer then reads) and
launches `all_reduce(AVG)` asynchronously.  (Round 1 pre-set `p.grad` to the zeroed views so that autograd accumulated in
place "without a copy" -- but AccumulateGrad then runs one `add_` kernel per parameter, 826 launches = 1.85 ms of a
22 ms step, measured with the reducer forced on ONE GPU; that, not NCCL, was the whole 1 -> 8 GPU scaling loss.) ProcessGroupNCCL runs it on its own stream after an event on the compute
stream, and `finish()` makes the compute stream wait for all of them before the optimizer reads the gradients.
Everything here is stream-ordered (no host sync), so a whole step -- forward, backward, the all-reduces and the
optimizer -- can be captured into one CUDA graph (rdst_b200/train.py).

torch's DistributedDataParallel also works on the module (it sees the same link-by-link readiness); this class is
the graph-capturable, copy-free alternative.  The reference trains on a single GPU (models/trans_sr_trainer.py), so
there is no reference interface to mirror here.
"""
import torch
import torch.distributed as dist


def rdst_link_of(name):
    """Bucket key of a parameter name of RDSTSR / RDSTSR_N ('head' | 'body.<i>' | 'tail'; state_dict layout) or
    of SwinIR ('head' | 'layers.<i>' | 'tail'): one bucket per link of the autograd chain."""
    parts = name.split(".")
    if parts[0] in ("body", "layers"):
        return parts[0] + "." + parts[1]
    if parts[0] in ("head", "conv_first", "patch_embed"):
        return "head"
    return "tail"


class BucketedAllReduce:
    def __init__(self, model, process_group=None, bucket_of=rdst_link_of, average=True, allow_unused=False):
        self.group = process_group
        self.average = average
        self.allow_unused = allow_unused    # parameters the forward never uses (RDSTSR_N's norm / conv_after_body): their
                                            # bucket is reduced at finish() with zeros in the unused slots
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        groups = {}
        for name, p in model.named_parameters():
            if p.requires_grad:
                groups.setdefault(bucket_of(name), []).append(p)
        self.buckets = []
        self._hooks = []
        for key, params in groups.items():
            n = sum(p.numel() for p in params)
            flat = torch.zeros(n, dtype=params[0].dtype, device=params[0].device)
            views, off = [], 0
            for p in params:
                views.append(flat[off:off + p.numel()].view_as(p))
                off += p.numel()
            b = dict(key=key, params=params, flat=flat, views=views, pending=len(params), fired=False)
            self.buckets.append(b)
            for p in params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._make_hook(b)))
        self._works = []
        self.launch_order = []          # bucket keys in the order their all-reduce was launched (last step)
        self.begin_step()

    def _make_hook(self, b):
        def hook(_param):
            b["pending"] -= 1
            if b["pending"] == 0:
                b["pending"] = len(b["params"])      # re-armed for the next backward, with or without begin_step()
                self._gather(b)
                self._launch(b)
        return hook

    @torch.no_grad()
    def _gather(self, b):
        """The bucket's gradients (fresh tensors autograd has just assigned to p.grad) -> the flat buffer, one multi-tensor
        copy; afterwards p.grad ARE the views (parameters without a gradient keep zeros)."""
        src, dst = [], []
        for p, v in zip(b["params"], b["views"]):
            if p.grad is not None and p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(b["params"], b["views"]):
            p.grad = v

    def _launch(self, b):
        b["fired"] = True
        self.launch_order.append(b["key"])
        if self.world == 1:
            return
        if self.average and dist.get_backend(self.group) == "nccl":
            self._works.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        else:
            if self.average:
                b["flat"].div_(self.world)
            self._works.append(dist.all_reduce(b["flat"], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def begin_step(self):
        """Replaces optimizer.zero_grad(): zero the flat buffers and drop the gradients (autograd then ASSIGNS the new
        ones instead of launching an add per parameter)."""
        self._works, self.launch_order = [], []
        for b in self.buckets:
            b["flat"].zero_()
            b["pending"], b["fired"] = len(b["params"]), False
            for p in b["params"]:
                p.grad = None

    def finish(self):
        """Call after backward(): the current stream waits for every bucket's all-reduce."""
        missing = [b for b in self.buckets if not b["fired"]]
        if missing and not self.allow_unused:
            raise RuntimeError(f"rdst_b200.ddp: buckets {[b['key'] for b in missing]} received no gradient in this backward "
                               "pass (pass allow_unused=True if the model has parameters its forward never uses)")
        for b in missing:                   # same set on every rank: the collective stays matched
            self._gather(b)
            self._launch(b)
        for w in self._works:
            w.wait()
        self._works = []
        for b in self.buckets:              # a loop that calls optimizer.zero_grad() instead of begin_step() keeps working: the
            b["fired"] = False              # next backward assigns fresh gradients and the hooks gather them again
            b["pending"] = len(b["params"])

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []
