"""Clip sharding across the GPUs of one box (one process per GPU, torch.distributed).

The lifting path shards on independent units -- clips (batch axis): inference needs NO
collective on the data path.  The only exchange step of the path is the
training config's gradient all-reduce: one NCCL all-reduce over a flat fp32 buffer
(27.66 MB at 27f/17j/128ch), which replaces the reference's single-process
nn.DataParallel reduce (trainval.py:56-61).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced split of n_items over `world` ranks: first (n % world) ranks get
    one extra.  Returns (start, stop)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError('bad rank/world %r/%r' % (rank, world))
    base, extra = divmod(int(n_items), world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def forward_sharded(fn, x_global, rank=None, world=None, gather=False, group=None):
    """Run `fn` (e.g. a SpatioTemporalModel on this rank's GPU) on this rank's clips of
    x_global (B,T,J,F).  No collective unless gather=True, which all-gathers the per-rank
    outputs into the global (B,T_out,J,3) order (ragged shards are padded to the largest)."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    s, e = shard_range(x_global.shape[0], rank, world)
    y = fn(x_global[s:e].contiguous()) if e > s else None
    if not gather or world == 1:
        return y
    sizes = [shard_range(x_global.shape[0], r, world) for r in range(world)]
    maxn = max(b - a for a, b in sizes)
    # every rank must know the trailing shape; broadcast it from the first non-empty rank
    shape = torch.zeros(8, dtype=torch.int64, device=x_global.device)
    if y is not None:
        shape[0] = y.dim() - 1
        shape[1:y.dim()] = torch.tensor(list(y.shape[1:]), dtype=torch.int64)
    dist.all_reduce(shape, op=dist.ReduceOp.MAX, group=group)
    trail = [int(v) for v in shape[1:1 + int(shape[0])]]
    dtype = y.dtype if y is not None else torch.float32
    pad = torch.zeros([maxn] + trail, dtype=dtype, device=x_global.device)
    if y is not None:
        pad[:y.shape[0]] = y
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:b - a] for o, (a, b) in zip(outs, sizes)], dim=0)


class FlatGradBuffer(object):
    """One contiguous fp32 buffer aliasing every parameter's .grad, so that the gradient
    exchange is a single all-reduce (sum, then 1/world) instead of ~200 small ones."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device('cpu')
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.div_(dist.get_world_size(group))
        return self.flat
