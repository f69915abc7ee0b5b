"""Stream (file) sharding across ranks — the only multi-GPU structure of the path.

One process per GPU; every rank owns a contiguous block of global stream ids and runs them with no
data-path collective. The only collective is the max-over-ranks of the timer (and, for reporting,
a sum of byte counts), issued through torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""


def stream_block(n_streams_total, world_size, rank):
    """Global stream ids owned by `rank`: contiguous, sizes differ by at most one."""
    base, extra = divmod(n_streams_total, world_size)
    lo = rank * base + min(rank, extra)
    return list(range(lo, lo + base + (1 if rank < extra else 0)))


def reduce_timing(dist, device, seconds, n_bytes):
    """max over ranks of the elapsed time, sum over ranks of the bytes processed."""
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    b = torch.tensor([float(n_bytes)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(b, op=dist.ReduceOp.SUM)
    return float(t.item()), float(b.item())
