"""Multi-GPU plumbing: pairs are independent through the whole path, so ranks own contiguous blocks
of pairs and exchange nothing until the end, where the per-pair [12] result rows (9 H entries, corner
error, inlier count, status) are all-gathered once."""
import torch
import torch.distributed as dist


def shard_range(global_pairs, rank, world):
    """Contiguous block of pairs for ``rank``; sizes differ by at most one."""
    base, rem = divmod(global_pairs, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_results(local_rows, global_pairs=None):
    """all_gather of ``[B_local, 12]`` rows -> ``[B_global, 12]`` on every rank (ragged shards allowed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_rows
    world = dist.get_world_size()
    n = torch.tensor([local_rows.shape[0]], device=local_rows.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    mx = max(sizes)
    pad = torch.zeros((mx, local_rows.shape[1]), device=local_rows.device, dtype=local_rows.dtype)
    pad[: local_rows.shape[0]] = local_rows
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    out = torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)
    if global_pairs is not None:
        assert out.shape[0] == global_pairs
    return out


def gather_equal(local):
    """ONE collective for equally sized shards: ``[..., 12]`` rows of every rank -> ``[world, ...]`` on every rank
    (``all_gather_into_tensor``: no per-rank copy kernels).  The benchmark calls it once after its K steps (SURVEY.md
    8(e): "a single gather at the very end"), not once per step."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local[None]
    out = torch.empty((dist.get_world_size(),) + tuple(local.shape), device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out
