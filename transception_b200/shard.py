"""Data-parallel plumbing for the hot path: images shard across ranks, one process per GPU.

The forward needs no data-path collective — rank ``r`` of ``W`` owns images ``[r*per, (r+1)*per)`` of the global
batch (reference ``trainer.py:86``: ``batch_size * n_gpu``) — so the only cross-rank traffic here is the
max-over-ranks reduction of device-measured times used by ``bench.py``.  Works on any ``torch.distributed``
backend (NCCL on the GPU box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(global_batch, rank, world):
    """Half-open image range of ``rank``; the global batch must divide evenly (weak scaling: fixed per-rank batch)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    if global_batch % world:
        raise ValueError("global batch %d does not divide over %d ranks" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def shard_batch(x, rank, world):
    """The slice of a global batch tensor ``[B, ...]`` owned by ``rank``."""
    lo, hi = shard_range(x.shape[0], rank, world)
    return x[lo:hi]


def max_over_ranks(values, device="cpu"):
    """Element-wise MAX of a list of floats over all ranks (identity when not initialised / world 1)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_rows(x, world):
    """All-gather equal-sized shards back into the global batch order (used by tests / multi-GPU inference)."""
    if not (dist.is_available() and dist.is_initialized()) or world == 1:
        return x
    parts = [torch.empty_like(x) for _ in range(world)]
    dist.all_gather(parts, x.contiguous())
    return torch.cat(parts, 0)


class GradBucket:
    """One flat fp32 bucket holding the gradients of the USED parameters, averaged over the ranks with a single
    all-reduce per step (1 217 tensors / 38.1 M elements for MSTransception; the 332 parameters the forward
    never touches keep ``grad is None`` on every rank, as under the reference's nn.DataParallel + SGD).  BatchNorm statistics
    stay per rank, like the reference's replicas."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]

    def allreduce(self, group=None):
        """Average ``p.grad`` over the ranks in place; returns the number of elements reduced."""
        grads = [p.grad for p in self.params if p.grad is not None]
        if not grads:
            return 0
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        n = sum(g.numel() for g in grads)
        if world == 1:
            return n
        flat = torch.cat([g.reshape(-1) for g in grads])
        if dist.get_backend(group) == "nccl":
            dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)      # averaged inside the collective
        else:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)      # gloo has no AVG
            flat.mul_(1.0 / world)
        views, off = [], 0
        for g in grads:
            views.append(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        torch._foreach_copy_(grads, views)        # a handful of multi-tensor copies instead of one kernel per parameter
        return n
