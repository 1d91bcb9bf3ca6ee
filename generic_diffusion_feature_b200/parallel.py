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
    """Gather per-rank result rows (argmax indices of this rank's image pairs, resized feature stacks, ...; shape
    [n_local, ...]) on rank 0. `counts`: rows per rank (defaults to equal). Returns the concatenated tensor on rank 0,
    None elsewhere.
    Point-to-point: every non-root rank sends exactly its own rows to rank 0 (NCCL send / recv over NVLink, batched
    into one group) - no padding and nothing delivered to ranks that do not need it (an all_gather would move
    world x the bytes). Rank 0 receives straight into slices of the output tensor."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return tensor
    world, rank = dist.get_world_size(), dist.get_rank()
    if counts is None:
        counts = [tensor.shape[0]] * world
    tensor = tensor.contiguous()
    if rank != 0:
        if counts[rank] > 0:
            dist.send(tensor[:counts[rank]], dst=0)
        return None
    out = torch.empty((sum(counts),) + tuple(tensor.shape[1:]), dtype=tensor.dtype, device=tensor.device)
    out[:counts[0]] = tensor[:counts[0]]
    ops, off = [], counts[0]
    for r in range(1, world):
        if counts[r] > 0:
            ops.append(dist.P2POp(dist.irecv, out[off:off + counts[r]], r))
        off += counts[r]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def gather_bytes(counts, row_bytes):
    """Bytes that cross the interconnect in gather_to_rank0 (everything but rank 0's own rows)."""
    return sum(counts[1:]) * row_bytes
