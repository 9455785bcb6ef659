"""Multi-GPU: options are independent, so the batch is cut into contiguous slices, one per rank, with no collective
on the data path; an optional all_gather of the 8 B/option prices closes the job."""
import torch


def shard_bounds(N, world_size, rank):
    """Contiguous slice [lo, hi) of rank `rank`: sizes differ by at most one, lower ranks take the remainder."""
    base, rem = divmod(int(N), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def price_batch_sharded(r, q, sigma, K, T, S, is_put, t=None, cfg=None, gather=True, group=None):
    """Every rank holds (or can index) the full input arrays; it prices its own slice on its own GPU and, if
    `gather`, all ranks receive the full price vector (one all_gather of padded slices over NCCL/NVLink, or gloo
    when the tensors live on the CPU in tests).  Returns (prices, (lo, hi))."""
    import torch.distributed as dist
    from . import batch
    cfg = cfg or batch.AloConfig()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N = len(r)
    lo, hi = shard_bounds(N, world, rank)
    sl = slice(lo, hi)
    out = batch.price_batch(r[sl], q[sl], sigma[sl], K[sl], T[sl], S[sl], is_put[sl], t=None if t is None else t[sl],
                            cfg=cfg, want_boundary=False)
    local = out["price"]
    if not gather or world == 1:
        return local, (lo, hi)
    return gather_slices(local, N, world, group), (lo, hi)


def gather_slices(local, N, world, group=None, unit=1):
    """all_gather of per-rank slices of unequal length (padded to the longest) back into one length-N vector.
    unit: slices are cut in whole rows of `unit` elements (N = rows * unit), as ladder_batch_sharded cuts scenarios."""
    import torch.distributed as dist
    rows = N // unit
    longest = -(-rows // world) * unit
    pad = torch.zeros(longest, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.empty(N, dtype=local.dtype, device=local.device)
    for rk in range(world):
        lo, hi = shard_bounds(rows, world, rk)
        full[lo * unit:hi * unit] = parts[rk][:(hi - lo) * unit]
    return full


def interleaved_rows(G, world_size, rank):
    """Rows g = rank, rank + W, rank + 2W, ... of a G-row problem (where cost correlates with the row index — the
    sweep count and the number of recorded iterates of a (T, sigma) scenario grow with T — deal the rows round-robin instead of
    cutting contiguous blocks, so every rank sees the same mix)."""
    return slice(int(rank), int(G), int(world_size))


def gather_interleaved(local, G, world, group=None, unit=1):
    """all_gather of per-rank row sets dealt round-robin (interleaved_rows) back into one G x unit matrix, flattened."""
    import torch.distributed as dist
    longest = -(-G // world) * unit
    pad = torch.zeros(longest, dtype=local.dtype, device=local.device)
    pad[:local.numel()] = local.reshape(-1)
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    full = torch.empty((G, unit), dtype=local.dtype, device=local.device)
    for rk in range(world):
        cnt = len(range(rk, G, world))
        full[rk::world] = parts[rk][:cnt * unit].reshape(cnt, unit)
    return full.reshape(-1)


def ladder_batch_sharded(r, q, sigma, T, S, is_put, K, cfg=None, gather=False, group=None):
    """Strike ladders (batch.ladder_batch) with the SCENARIOS dealt round-robin across ranks: rank g prices rows g, g + W, ...
    of the G x nK result — the boundary solves are split too, nothing is replicated and nothing is exchanged.  With `gather`
    every rank receives the full G x nK matrix (one all_gather).  Returns (prices, rows) with rows = the slice of row indices
    this rank priced."""
    import torch.distributed as dist
    from . import batch
    cfg = cfg or batch.AloConfig()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    G = len(r)
    sl = interleaved_rows(G, world, rank)

    def take(x):
        return x[sl].contiguous() if isinstance(x, torch.Tensor) else x[sl]
    local = batch.ladder_batch(take(r), take(q), take(sigma), take(T), take(S), take(is_put), K, cfg=cfg)
    if not gather or world == 1:
        return local, sl
    nK = local.shape[1]
    return gather_interleaved(local, G, world, group, unit=nK).reshape(G, nK), sl
