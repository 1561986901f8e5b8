"""Ray sharding over the GPUs of one box: rays are independent, so a global batch is cut into contiguous
per-rank shards, every rank normalises its loss by the GLOBAL ray count, and one all-reduce (sum) of the flat gradient
vector (two squared-error sums ride along) makes every rank apply the identical Adam update.  RNG counters are keyed
by the global ray index (shard offset), so results do not depend on the number of ranks.  No CUDA dependency here:
the same functions are exercised with the gloo backend in tests/test_distributed_cpu.py."""
import torch


def shard_bounds(n_total: int, world_size: int, rank: int):
    """Contiguous shard [lo, hi) of rank `rank`; the last shards may be shorter or empty when world_size does not divide."""
    per = (n_total + world_size - 1) // world_size
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


def allreduce_sum_(flat: torch.Tensor, group=None, async_op: bool = False):
    """In-place sum over ranks of (a contiguous slice of) the flat [sq_err_c, sq_err_f, 0, 0 | grads_coarse | grads_fine]
    buffer.  async_op=True returns the work handle (None when there is nothing to reduce): the fine-network slice is
    reduced while the coarse backward still runs."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            return work
    return None if async_op else flat


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate the ranks' contiguous row blocks (``shard_bounds`` of ``n_total`` rows; the last blocks may be shorter
    or empty) into the full (n_total, ...) tensor on every rank.  Blocks are padded to a common length so that the plain
    equal-size all-gather of every backend applies."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return local
    per = (n_total + world - 1) // world
    padded = local if local.shape[0] == per else torch.cat(
        [local, local.new_zeros((per - local.shape[0],) + tuple(local.shape[1:]))])
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous(), group=group)
    out = []
    for r, part in enumerate(parts):
        lo, hi = shard_bounds(n_total, world, r)
        out.append(part[:hi - lo])
    return torch.cat(out)
