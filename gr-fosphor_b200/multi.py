"""Multi-GPU plumbing (one process per GPU, torch.distributed).

The hot path shards by CHANNEL (option A): every channel
is an independent fosphor instance with its own histogram / live / max-hold /
waterfall state - in the reference that is simply several sink blocks, whose
init is serialised by a static mutex (lib/base_sink_c_impl.cc:46,97).  There is
no data-path collective; the only exchange is the reduced max-hold trace
(BASELINE.json configs[3]): one all-reduce(MAX) over N floats.
"""


def channels_for_rank(rank, world, n_channels):
    """Static round-robin partition of independent channels over ranks."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return [c for c in range(n_channels) if c % world == rank]


def reduce_maxhold(dist, trace):
    """In-place max over ranks of a max-hold trace (torch tensor, N floats).
    `dist` is torch.distributed (nccl on GPUs, gloo in the CPU tests)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(trace, op=dist.ReduceOp.MAX)
    return trace


def max_over_ranks(dist, value, device=None):
    """Timing helper: max of a python float over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
