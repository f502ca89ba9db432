"""Multi-GPU partitioning of the sampling path: independent joint samples, no collective on the data path.

A joint sample = one camera row + one lidar row (+ their CFG duplicates); cross-modal attention pairs rows 2i and
2i+1 only (attention.py:246-263), so shards are cut between samples, never inside one.  Every sample draws its noise
from its own generator seeded by (base seed, global sample index), so results do not depend on how many GPUs the
batch was cut into.  The reference has no multi-GPU inference path (scripts/inference_test_bench.py is single-process);
this is the process-per-GPU equivalent of running that script on slices of the dataset.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_samples, world, rank):
    """Contiguous, balanced [lo, hi) range of joint-sample indices owned by `rank` (first n % world ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(n_samples, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(t, world, rank, rows_per_sample=2):
    """Rows of an interleaved [n_samples * rows_per_sample, ...] tensor that belong to `rank`."""
    n = t.shape[0] // rows_per_sample
    if n * rows_per_sample != t.shape[0]:
        raise ValueError("leading dim %d is not a multiple of rows_per_sample=%d" % (t.shape[0], rows_per_sample))
    lo, hi = shard_bounds(n, world, rank)
    return t[lo * rows_per_sample:hi * rows_per_sample]


def sample_noise(shape_per_row, lo, hi, base_seed=0, rows_per_sample=2, device="cpu", dtype=torch.float32):
    """x_T rows for joint samples lo..hi-1: sample i uses torch.Generator().manual_seed(base_seed * 1000003 + i)."""
    rows = []
    for i in range(lo, hi):
        g = torch.Generator(device="cpu").manual_seed(base_seed * 1000003 + i)
        rows.append(torch.randn((rows_per_sample,) + tuple(shape_per_row), generator=g, dtype=dtype))
    if not rows:
        return torch.empty((0,) + tuple(shape_per_row), dtype=dtype, device=device)
    return torch.cat(rows).to(device)


def gather_samples(local, n_samples, rows_per_sample=2, group=None):
    """All-gathers the per-rank result rows into the full [n_samples * rows_per_sample, ...] tensor on every rank
    (host-side convenience for writing results from one rank; NOT part of the timed data path)."""
    if not dist.is_available() or not dist.is_initialized():
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_bounds(n_samples, world, r) for r in range(world)]
    biggest = max(hi - lo for lo, hi in counts) * rows_per_sample
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:(hi - lo) * rows_per_sample] for b, (lo, hi) in zip(bufs, counts)])
