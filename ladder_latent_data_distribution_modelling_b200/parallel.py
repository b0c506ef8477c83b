"""Multi-GPU plumbing of the hot path (one process per GPU, torch.distributed).

Two decompositions: batch-sharded data parallelism (the engine all-reduces its 12
batch sums and each group's flat gradient) and the component-sharded hyper-prior: rank r holds
components [lo_r, hi_r), evaluates the (m, s, unnormalised g) partial with the fused kernel, the
partials are all-gathered and combined.  `ShardedMixture` is the production form: the kernel
writes ONE packed partial [N, 2 + D], the ranks trade ONE `all_gather_into_tensor` into a
preallocated [P, N, 2 + D] buffer, one combine kernel finishes, and the whole call (kernel ->
NCCL -> combine) is replayed as a CUDA graph.  `sharded_mixture_logprob` keeps the three-tensor
list form (any backend; the gloo tests on CPU drive its host logic).
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


class ShardedMixture:
    """Component-sharded log p(t) (+ d log p / d t) for a FIXED query count N: one kernel, one all-gather, one combine.

    Every rank passes the full packed table and the same queries [N, D]; the answer is complete on every rank.  Buffers are
    allocated once; with `graph=True` (default on CUDA) the three launches are captured -- the NCCL collective included -- and
    each call is a copy of the queries into the static input plus one graph replay."""

    def __init__(self, table, N, group=None, want_grad=False, graph=True):
        from . import ops
        self.ops = ops
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.shard = table.shard(self.rank, self.world)
        self.N, self.D, self.want_grad = int(N), table.D, bool(want_grad)
        dev = table.table.device
        W = 2 + self.D if want_grad else 2
        self.t = torch.empty(self.N, self.D, device=dev)
        self.pack = torch.empty(self.N, W, device=dev)
        self.parts = torch.empty(self.world, self.N, W, device=dev) if self.world > 1 else self.pack.view(1, self.N, W)
        self.logp = torch.empty(self.N, device=dev)
        self.grad = torch.empty(self.N, self.D, device=dev) if want_grad else None
        self._graph = None
        self.use_graph = bool(graph) and dev.type == 'cuda'

    def release(self):
        """Drop the captured graph (it holds an NCCL kernel: do this before destroying the process group)."""
        self._graph = None

    def _launch(self):
        ops = self.ops
        ops.mixture_logprob_packed(self.t, self.shard, self.pack, self.want_grad)
        if self.world > 1:
            dist.all_gather_into_tensor(self.parts, self.pack, group=self.group)
        ops.mixture_combine_packed(self.parts, self.D, self.want_grad, self.logp, self.grad)

    def __call__(self, t):
        """t [N, D] -> logp [N] (, grad [N, D]); the returned tensors are the object's static outputs."""
        self.t.copy_(t)
        if not self.use_graph:
            self._launch()
        else:
            if self._graph is None:
                side = torch.cuda.Stream()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    self._launch()                     # warm-up: workspaces, NCCL channels
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                try:
                    with torch.cuda.graph(g, capture_error_mode='thread_local' if self.world > 1 else 'global'):
                        self._launch()
                    self._graph = g
                except Exception as e:                 # noqa: BLE001
                    import sys
                    print('[ladder] ShardedMixture: graph capture failed (%s); launching eagerly' % type(e).__name__,
                          file=sys.stderr)
                    self.use_graph = False
                    torch.cuda.synchronize()
                    self._launch()
                    return (self.logp, self.grad) if self.want_grad else self.logp
            self._graph.replay()
        return (self.logp, self.grad) if self.want_grad else self.logp
