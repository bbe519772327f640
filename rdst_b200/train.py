"""Whole training step as ONE CUDA graph: forward, loss, backward, gradient all-reduce and optimizer.

The eager training step of the RDST module enqueues ~1200 kernels of librdst_b200 plus a few thousand small torch ops
(the differentiable weight packing and its backward); at the cfg4 batch (32 x 1x24x24) the host needs longer to
enqueue them than the B200 needs to run them.  `GraphedTrainStep` captures one step on static input/target buffers
and replays it with a single launch; data-parallel runs pass a `rdst_b200.ddp.BucketedAllReduce`, whose NCCL
all-reduces are captured on ProcessGroupNCCL's stream inside the same graph and overlap with the backward kernels.

The reference's step being replaced: TransSRTrainer.train's inner loop (models/trans_sr_trainer.py:140-160:
zero_grad -> G(lr) -> loss -> backward -> optimizer.step), L1 loss + Adam per RDST_E1_OASIS_example_SRx4.ini:130-135.
"""
import torch


class GraphedTrainStep:
    """step = GraphedTrainStep(model, optimizer, x_example, y_example); loss = step(x, y)

    * `optimizer` must be capturable (torch.optim.Adam(..., capturable=True)).
    * `reducer`: optional rdst_b200.ddp.BucketedAllReduce for data-parallel training.
    * The warm-up / capture steps run on the example batch; parameters and optimizer state are restored afterwards,
      so the first call of the object is the first real optimisation step.
    * Returned loss is a device tensor that is overwritten by the next call (clone it to keep it)."""

    def __init__(self, model, optimizer, x_example, y_example, loss_fn=torch.nn.functional.l1_loss, reducer=None,
                 warmup=3):
        for g in optimizer.param_groups:
            if not g.get("capturable", False):
                raise ValueError("rdst_b200.train: the optimizer must be created with capturable=True to be "
                                 "captured into a CUDA graph")
        self.model, self.opt, self.loss_fn, self.reducer = model, optimizer, loss_fn, reducer
        self.x = x_example.detach().clone()
        self.y = y_example.detach().clone()
        params = [p for g in optimizer.param_groups for p in g["params"]]
        saved_p = [p.detach().clone() for p in params]
        saved_s = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in optimizer.state[p].items()}
                   for p in params if p in optimizer.state and optimizer.state[p]}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._one_step()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        if reducer is None:
            optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._one_step()
        # undo the warm-up / capture updates (in place: the graph holds these addresses)
        with torch.no_grad():
            for p, s in zip(params, saved_p):
                p.copy_(s)
            for p in params:
                st = optimizer.state.get(p, {})
                old = saved_s.get(id(p))
                for k, v in st.items():
                    if torch.is_tensor(v):
                        if old is not None and k in old:
                            v.copy_(old[k])
                        else:
                            v.zero_()
        torch.cuda.synchronize()

    def _one_step(self):
        if self.reducer is not None:
            self.reducer.begin_step()
        else:
            self.opt.zero_grad(set_to_none=True)
        loss = self.loss_fn(self.model(self.x), self.y)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish()
        self.opt.step()
        return loss.detach()

    def __call__(self, x, y):
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)
        self.graph.replay()
        return self.loss
