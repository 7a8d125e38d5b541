"""Multi-GPU plumbing: one process per GPU, tile-sharded like exec::distrib (src/exec/distrib/master.rs:88-120),
film combined by SUM like film::Image::add_pixels (src/film/image.rs:21-50).

The reference partitions the Morton-sorted 8x8 block list into contiguous ranges of floor(B / W) blocks, the last
worker taking the remainder (master.rs:91-93, 218-224); `shard_blocks` reproduces that. The exchange step is one
reduce of the RGBW film (torch.distributed: NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_blocks(n_blocks, rank, world):
    """(block_start, block_count) of `rank`, exactly master.rs:91-93 + 218-224."""
    per = n_blocks // world
    start = rank * per
    count = per if rank != world - 1 else n_blocks - start
    return start, count


def shard_is_empty(shard):
    """True when `shard_blocks` gave this rank nothing (n_blocks < world). Such a rank must SKIP the render call:
    in the ABI (and in block_queue.rs:39-41) block_count == 0 selects ALL blocks, which would over-count the film."""
    return shard[1] == 0


def shard_interleaved(rank, world, chunk=32):
    """Load-balanced alternative: rank r takes every world-th chunk of `chunk` consecutive blocks of the Morton list
    (render-cfg fields shard_index / shard_count / shard_chunk). Same union, same summed film; image regions of very
    different cost (e.g. rays that leave the scene) are spread over all ranks instead of landing on one."""
    return dict(shard_index=rank, shard_count=world, shard_chunk=chunk) if world > 1 else {}


def reduce_film(film, dst=0, group=None):
    """Sum the per-rank RGBW films into rank `dst` (in place). `film` is a torch tensor on the rank's device."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return film
    dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film


def max_over_ranks(value, device):
    """Timing rule: the job time is the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(values, device):
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]
