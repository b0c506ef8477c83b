"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

Two decompositions: batch-sharded data parallelism (the engine all-reduces its 12
batch sums and each group's flat gradient) and the component-sharded hyper-prior: rank r holds
components [lo_r, hi_r), evaluates the (m, s, unnormalised g) partial with the fused kernel, the
partials are all-gathered and combined by `ladder_mixture_combine`.  The helpers here are device
agnostic so the host-side logic is testable with gloo on CPU.
"""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous, balanced-by-ceil partition of range(n): the shard of `rank`."""
    per = -(-n // world)
    return min(rank * per, n), min((rank + 1) * per, n)


def allreduce_sum_(t, group=None):
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_partials(m, s, g=None, group=None):
    """Stack every rank's (m [N], s [N], g [N, D]) on a new leading axis, rank-major."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return m[None], s[None], (g[None] if g is not None else None)

    def gather(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous(), group=group)
        return torch.stack(out)
    return gather(m), gather(s), (gather(g) if g is not None else None)


def sharded_mixture_logprob(t, table, group=None, want_grad=False):
    """log p(t) (and d log p / d t) with the mixture's components sharded over the ranks of `group`.
    Every rank passes the FULL packed table and the same queries; returns the full answer on every rank."""
    from . import ops
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    part = ops.mixture_logprob(t, table.shard(rank, world), want_grad=want_grad, partial=True)
    m, s, g = allgather_partials(part[0], part[1], part[2] if want_grad else None, group)
    return ops.mixture_combine(m, s, g)
