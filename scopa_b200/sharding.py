"""How one MCCFR iteration's traversals are split over ranks (one process per GPU).

Traversal ids are global, so the union of all ranks' work -- and therefore the summed delta -- does
not depend on the number of ranks.  The only exchange is one all-reduce(sum) of the slot-aligned
delta buffer per iteration; slot alignment holds because every rank numbers the
infosets by the same deterministic tree enumeration.
"""


def shard_bounds(total, rank, world):
    """Contiguous share [lo, lo + n) of `total` traversal ids for `rank` of `world`."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    share = (total + world - 1) // world
    lo = min(total, rank * share)
    hi = min(total, lo + share)
    return lo, hi - lo


def allreduce_delta(delta, group=None):
    """Sum the delta buffer (torch tensor, any backend: nccl on GPUs, gloo in the CPU tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=group)
    return delta
