"""Batched, slice-sharded inference driver.

The reference tester feeds one LR slice per forward call and synchronises device->host after each
(models/trans_sr_tester.py:124-166, models/basic_tester.py:104-115).  Slices never interact (attention is per 8x8
window, convolutions per image), so this driver
  * splits the slice axis of a volume into contiguous ranges, one per rank (no collective on the data path),
  * runs each rank's range in large batches through the drop-in module,
  * keeps host<->device copies asynchronous on pinned buffers.
"""
import torch


def shard_range(n_items, world_size, rank):
    """Contiguous, balanced [begin, end) range of `n_items` for `rank` (first n%world ranks get one extra)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


@torch.no_grad()
def super_resolve_slices(model, lr_slices, batch_size=176, out=None, use_graph=False):
    """lr_slices: (N,1,H,W) float tensor on the HOST (ideally pinned) or on the model's device.
    Returns (N,1,sH,sW) on the same side as the input.  One H2D + one D2H per batch, both asynchronous.
    use_graph: replay one captured CUDA graph per batch shape (`GraphedRDST`, cached on the model) instead of ~110 eager
    launches -- what a rank wants when its share of a volume is a few dozen slices (launch-latency bound)."""
    dev = next(model.parameters()).device
    n = lr_slices.shape[0]
    s = model.sr_scale
    on_host = not lr_slices.is_cuda
    if out is None:
        out = torch.empty(n, 1, lr_slices.shape[2] * s, lr_slices.shape[3] * s, dtype=torch.float32,
                          device=lr_slices.device, pin_memory=on_host and torch.cuda.is_available())
    # pinned host output: the reconstruction kernel writes the HR slices straight into it (no separate D2H copy)
    direct = (on_host and out.is_pinned() and out.dtype == torch.float32 and out.is_contiguous() and
              getattr(model, "_exec", None) is not None and "out" in model.forward.__code__.co_varnames)
    graphed = None
    if use_graph:
        graphed = getattr(model, "_graphed_forward", None)
        if graphed is None:
            graphed = GraphedRDST(model)
            object.__setattr__(model, "_graphed_forward", graphed)      # plain attribute: not a submodule, not in the state_dict
    for b0 in range(0, n, batch_size):
        x = lr_slices[b0:b0 + batch_size]
        if on_host:
            x = x.to(dev, non_blocking=True)
        if graphed is not None:
            out[b0:b0 + batch_size].copy_(graphed(x, clone=False), non_blocking=True)
        elif direct:
            model(x, out=out[b0:b0 + batch_size])
        else:
            out[b0:b0 + batch_size].copy_(model(x), non_blocking=True)
    if on_host and torch.cuda.is_available():
        torch.cuda.current_stream(dev).synchronize()
    return out


def super_resolve_volume_sharded(model, lr_volume, rank=0, world_size=1, batch_size=176, use_graph=None):
    """Each rank super-resolves its contiguous share of the slice axis; returns (begin, end, hr_slices).
    use_graph=None: graph replay when the share is smaller than one batch and the model sits on a GPU."""
    b, e = shard_range(lr_volume.shape[0], world_size, rank)
    if use_graph is None:
        use_graph = world_size > 1 and next(model.parameters()).is_cuda
    return b, e, super_resolve_slices(model, lr_volume[b:e], batch_size, use_graph=use_graph)


class GraphedRDST:
    """CUDA-graph replay of the drop-in forward for fixed input shapes (bf16 / fp32 inference).

    The reference tester calls the network once per LR slice (models/trans_sr_tester.py:146-152); at that size the
    forward is ~110 dependent kernel launches and is launch-latency bound.  All kernels of librdst_b200 are
    stream-ordered and allocation-free, so one forward per (B, H, W) is captured once and replayed on static buffers.
    Weight updates (load_state_dict, optimizer steps) invalidate the captured graphs.
    """

    def __init__(self, model):
        self.model = model
        self._graphs = {}
        self._wkey = None

    def _weights_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.model.parameters())

    @torch.no_grad()
    def __call__(self, x, clone=True):
        """clone=False returns the graph's static output buffer (valid until the next replay of the same shape)."""
        if not x.is_cuda:
            raise RuntimeError("GraphedRDST: input must be a CUDA tensor")
        wkey = self._weights_key()
        if wkey != self._wkey:
            self._graphs.clear()
            self._wkey = wkey
        key = (tuple(x.shape), x.dtype, str(x.device), self.model.precision)
        entry = self._graphs.get(key)
        if entry is None:
            static_in = x.clone()
            self.model(static_in)                               # builds weight cache + workspaces outside the capture
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                self.model(static_in)
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_out = self.model(static_in)
            # the graph addresses the executor's workspace buffers by raw pointer: hold the dict so that the executor's
            # bounded workspace cache can evict the shape without freeing memory this graph replays on
            held = getattr(getattr(self.model, "_exec", None), "last_workspace", None)
            entry = (graph, static_in, static_out, held)
            self._graphs[key] = entry
        graph, static_in, static_out, _ = entry
        static_in.copy_(x)
        graph.replay()
        return static_out.clone() if clone else static_out
