"""Multi-GPU plumbing of the extraction path: images (and image pairs) are independent, so the path shards
with NO data-path collective - weights are replicated, every rank extracts its own contiguous shard into its own
arena. torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests) is used only for (a) the barrier /
max-over-ranks timing of bench.py and (b) gathering small results (correspondence indices, resized stacks) on one
rank when a consumer wants them there. The reference has no distributed code on this path at all (SURVEY.md 2.1:
one FeatureExtractor per GPU driven from Python threads, aggregation_network.py:86-93)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced split of range(n_items): the first n_items % world ranks get one extra item."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cpu"):
    """Whole-job step time = max over ranks (every multi-GPU number is timed on the device, max over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_to_rank0(tensor, counts=None):
    """Gather per-rank result rows (e.g. argmax indices of this rank's image pairs, shape [n_local, ...]) on rank 0.
    `counts`: rows per rank (defaults to equal). Returns the concatenated tensor on rank 0, None elsewhere."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    world, rank = dist.get_world_size(), dist.get_rank()
    if counts is None:
        counts = [tensor.shape[0]] * world
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    pad[:tensor.shape[0]] = tensor
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != 0:
        return None
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)
