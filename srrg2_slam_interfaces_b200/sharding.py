"""Host-side sharding of the aligner's query (moving) cloud across ranks.

The moving cloud is split into contiguous index ranges, the fixed cloud is replicated; every rank
accumulates exact 64-bit fixed-point partial sums of H, b, chi and the counters, so a plain integer
all-reduce (NCCL on GPUs, gloo in the CPU tests) reproduces the single-rank result bit for bit."""


def shard_range(n, rank, world):
    """Contiguous [begin, end) of `n` items owned by `rank` out of `world` (sizes differ by <= 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def candidate_shard(n_candidates, rank, world):
    """Loop closing over several GPUs: the K candidate alignments of the brute-force detector are
    independent, so they are dealt to the ranks in contiguous runs (each rank batches its own run with
    srrg2b_closure_batch) and the per-candidate results are concatenated in rank order -- candidate order, as the
    reference's serial loop produces them.  No data-path collective."""
    begin, end = shard_range(n_candidates, rank, world)
    return list(range(begin, end))
