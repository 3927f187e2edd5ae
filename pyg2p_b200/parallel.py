"""Multi-GPU sharding of the interpolation path (one process per GPU, torch.distributed).

* intertable BUILD: the source index is replicated, target points are split into contiguous
  blocks (the reference's `num_of_splits` row chunks, scipy_interpolation_lib.py:588-607, become
  the rank partition); the only exchange is one all-gather of the finished table shards - and only
  when every rank needs the whole table (NCCL over NVLink on GPUs, gloo in the CPU tests).
* intertable APPLY: messages (step x ensemble member) are split by MEMBER so that the accumulation
  V[t] - V[t-D] of one member stays on one GPU; no collective at all.

Nothing here computes on the CPU: the functions take and return tensors on whatever device the
process group works on.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(total, rank, world_size, align=1):
    """[lo, hi) of `rank`'s contiguous block of `total` items; block starts are multiples of `align`
    (e.g. the number of columns, to split whole target rows).  Blocks differ by at most one unit."""
    units = -(-total // align)
    base, rem = divmod(units, world_size)
    lo_u = rank * base + min(rank, rem)
    hi_u = lo_u + base + (1 if rank < rem else 0)
    return min(lo_u * align, total), min(hi_u * align, total)


def members_of_rank(n_members, rank, world_size):
    """round-robin ensemble members -> rank (all steps of one member stay together)."""
    return list(range(rank, n_members, world_size))


def all_gather_table(idx_shard, coef_shard, m_total, align=1, group=None):
    """Assemble SoA table shards ([k, m_r] per rank, the contiguous target blocks of `shard_bounds(m_total, r,
    world, align)`) into the full [k, m_total] table on every rank.  The shard sizes follow from the arguments,
    so nothing but the table itself crosses the wire: with equal shards each neighbour row is gathered straight
    into its place (k collectives per array, no staging copy); ragged shards are padded to the largest."""
    rank, ws = world()
    if ws == 1:
        return idx_shard, coef_shard
    k, m_r = idx_shard.shape
    bounds = [shard_bounds(m_total, r, ws, align) for r in range(ws)]
    sizes = [hi - lo for lo, hi in bounds]
    if sizes[rank] != m_r:
        raise ValueError(f'rank {rank} holds {m_r} targets, its block has {sizes[rank]}')
    m_max = max(sizes)

    def gather(shard):
        if all(s == m_max for s in sizes):
            full = shard.new_empty((k, m_total))
            for kk in range(k):
                dist.all_gather_into_tensor(full[kk], shard[kk].contiguous(), group=group)
            return full
        pad = shard.new_zeros((k, m_max))
        pad[:, :m_r] = shard
        flat = shard.new_empty((ws * k, m_max))          # concatenation along dim 0 (the form gloo and nccl share)
        dist.all_gather_into_tensor(flat, pad, group=group)
        buf = flat.view(ws, k, m_max)
        return torch.cat([buf[r, :, :sizes[r]] for r in range(ws)], dim=1).contiguous()

    return gather(idx_shard), gather(coef_shard)


def max_over_ranks(value, device):
    """device-timed duration -> max over ranks (the contract for every multi-GPU number)."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_pairs(pair, device):
    """(hi, lo) of every rank, in rank order (the partial double-double sums of the ADW reference radius)."""
    rank, ws = world()
    if ws == 1:
        return [tuple(pair)]
    mine = torch.tensor([float(pair[0]), float(pair[1])], dtype=torch.float64, device=device)
    full = torch.empty((ws * 2,), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(full, mine)
    h = full.cpu().tolist()
    return [(h[2 * r], h[2 * r + 1]) for r in range(ws)]
