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


def allreduce_sum_(flat: torch.Tensor, group=None):
    """In-place sum over ranks of the flat [grads_coarse | grads_fine | sq_err_c | sq_err_f] buffer."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
