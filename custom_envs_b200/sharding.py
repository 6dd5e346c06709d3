"""Multi-GPU: envs shard by index, one process per GPU, no collective on the step path.

Envs never interact (reference vectorize/optvecenv.py:70-88 only concatenates their rows),
so rank r owns the contiguous env range ``shard_range(E, world, r)``, keeps a replica of
the data set and steps its own ``BatchedOptEnv``; the VecEnv row order of the whole job is
the concatenation of the shards.  The only exchange is an all-gather of small per-env
episode statistics (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(num_envs, world_size, rank):
    """Contiguous split of ``num_envs`` envs; the first ``num_envs % world_size`` ranks get
    one extra env.  Returns (first_env, count)."""
    base, extra = divmod(int(num_envs), int(world_size))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def shard_seeds(seeds, world_size, rank):
    """Seeds (one per global env index) of the envs this rank owns."""
    first, count = shard_range(len(seeds), world_size, rank)
    return list(seeds[first:first + count])


def gather_env_stats(local_stats, group=None):
    """All-gather per-env statistics ``[E_local, k]`` into ``[E_total, k]`` in global env order.
    Shards may differ in size by one env, so the payload is padded to the largest shard."""
    if not (dist.is_available() and dist.is_initialized()):
        return local_stats
    world = dist.get_world_size(group)
    local_stats = local_stats.contiguous()
    count = torch.tensor([local_stats.shape[0]], device=local_stats.device, dtype=torch.int64)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    counts = [int(c.item()) for c in counts]
    width = max(counts)
    padded = local_stats.new_zeros((width,) + tuple(local_stats.shape[1:]))
    padded[:local_stats.shape[0]] = local_stats
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([part[:n] for part, n in zip(parts, counts)], dim=0)
