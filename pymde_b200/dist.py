"""Edge-sharded multi-GPU solve: one process per GPU, edges split across ranks,
X replicated, and ONE all-reduce of [gradient | loss] per evaluation over NCCL/NVLink.

E and grad E are sums over edges (pymde/average_distortion.py:51,77-78), so each rank runs the
same fused kernel on its shard with the GLOBAL edge count as divisor; after the all-reduce
every rank holds bit-identical gradient and loss, the device-resident L-BFGS state is
replicated and (thanks to fixed-order reductions) takes identical decisions on every rank."""
import torch
import torch.distributed as dist

from . import _lib


def shard_range(p, rank, world):
    """Contiguous edge range [lo, hi) of `rank` (balanced to within one edge)."""
    base, rem = divmod(int(p), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _Wrap(object):
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 3}


def make_allreduce(device, group=None):
    """ctypes-callable (user, buf, count, stream) -> int doing an in-place NCCL sum on torch's
    current stream (the solver enqueues on that same stream)."""
    cache = {}

    def cb(user, buf, count, stream):
        try:
            key = (buf, count)
            t = cache.get(key)
            if t is None:
                with torch.cuda.device(device):
                    t = torch.as_tensor(_Wrap(buf, count), device=device)
                cache[key] = t
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
            return 0
        except Exception:  # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return _lib.MDE_E_INVALID

    return cb


def shard_mde(mde_cls, n_items, embedding_dim, edges, make_function, constraint, device, rank=None,
              world_size=None, group=None):
    """Build this rank's MDE over its edge shard.

    edges: full (p,2) int64 tensor (host or device); make_function(lo, hi) -> distortion function
    for edges[lo:hi].  Returns an MDE whose evaluations / embed() are global."""
    rank = dist.get_rank(group) if rank is None else rank
    world_size = dist.get_world_size(group) if world_size is None else world_size
    p = int(edges.shape[0])
    lo, hi = shard_range(p, rank, world_size)
    mde = mde_cls(n_items, embedding_dim, edges[lo:hi].to(device), make_function(lo, hi), constraint, device=device)
    if world_size > 1:
        mde.__dict__["_dist"] = {"rank": rank, "world_size": world_size, "p_total": p,
                                 "allreduce": make_allreduce(torch.device(device), group)}
    return mde
