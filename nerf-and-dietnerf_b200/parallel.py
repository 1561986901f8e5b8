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


class PeerExchange:
    """Gradient exchange over NVLink peer memory (``nerf_peer_barrier`` + ``nerf_peer_reduce_adam``): the flat gradient
    buffer lives in symmetric memory (torch.distributed._symmetric_memory: one allocation per GPU, mapped into every peer),
    every rank reads every peer's slice directly and applies Adam in the same kernel.  Two buffers alternate from step to
    step, so a buffer is rewritten two barriers after its last remote read.  Raises when the box has no peer access or the
    torch build has no symmetric memory: the caller then keeps the NCCL all-reduce."""

    def __init__(self, n_floats: int, device, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        group = dist.group.WORLD if group is None else group
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass                                            # newer torch enables it at rendezvous
        self.bufs, self.handles = [], []
        for _ in range(2):
            t = symm_mem.empty(n_floats, dtype=torch.float32, device=device)
            t.zero_()
            self.bufs.append(t)
            self.handles.append(symm_mem.rendezvous(t, group))
        self.pad = symm_mem.empty(256, dtype=torch.int32, device=device)
        self.pad.zero_()
        torch.cuda.synchronize(device)
        self.pad_handle = symm_mem.rendezvous(self.pad, group)
        dist.barrier(group)                                 # every pad is zero before anybody signals
        self.parity = 0
        self.sums = torch.zeros(4, dtype=torch.float32, device=device)

    @property
    def grads(self):
        return self.bufs[self.parity]

    def flip(self):
        self.parity ^= 1

    def barrier(self, epoch: int, slot: int):
        from ._lib import call
        call("nerf_peer_barrier", int(self.pad_handle.buffer_ptrs_dev), self.rank, self.world, int(epoch) & 0xFFFFFFFF, int(slot))

    def reduce_adam(self, params, offset, n, opt, state_offset, t, reduced=None):
        """Sum floats [offset, offset+n) of the current gradient buffer over the ranks and (``params`` given) apply step
        ``t`` of ``opt`` (Adam) to ``params`` with the moments at ``state_offset``."""
        from ._lib import call, ptr
        m = opt._m[state_offset:state_offset + n] if params is not None else None
        v = opt._v[state_offset:state_offset + n] if params is not None else None
        call("nerf_peer_reduce_adam", ptr(params), int(self.handles[self.parity].buffer_ptrs_dev), self.world, int(offset), int(n),
             ptr(m), ptr(v), opt.learning_rate, opt.beta_1, opt.beta_2, opt.epsilon, int(t), ptr(reduced))
