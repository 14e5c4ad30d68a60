"""Multi-GPU scale-out of the hot path: REPLICAS ONLY.

One frame's scan-to-map is a serial chain over a map that fits one GPU, and frame k+1 needs frame k's pose and map
(/root/reference/src/odomEstimationClass.cpp:229-282), so the natural unit is an independent sequence: one process per GPU,
one sequence per process, no collective in the frame loop.  torch.distributed is used only
to line the ranks up (barrier) and to combine the per-rank timings into the whole-job number.
"""
import torch
import torch.distributed as dist


N_SEQUENCES = 8     # BASELINE.json configs[4]: 8 independent sequences, seeds 3000..3007


def sequence_for_rank(rank, world_size):
    """Name of the synthetic sequence a rank processes.  The same family at every world size (a 64-ring street sequence of the
    configs[1] shape, seed 3000 + rank), so that the 1 / 2 / 4 / 8-GPU numbers compare like for like."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return f"cfg5.{rank % N_SEQUENCES}"


def sequences_for_rank(rank, world_size, total=N_SEQUENCES):
    """configs[4] as a fixed job of `total` sequences: the ones this rank runs (concurrently on its GPU) at this world size."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return [f"cfg5.{k}" for k in range(total) if k % world_size == rank]


def aggregate_throughput(frames_this_rank, elapsed_ms_this_rank, device=None):
    """Whole-job scans/s: frames of ALL ranks / MAX-over-ranks elapsed time.  Returns (scans_per_s, max_ms, total_frames)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([float(elapsed_ms_this_rank)], dtype=torch.float64, device=device)
        f = torch.tensor([float(frames_this_rank)], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
        max_ms, total = float(t[0]), float(f[0])
    else:
        max_ms, total = float(elapsed_ms_this_rank), float(frames_this_rank)
    return total / (max_ms * 1e-3), max_ms, int(total)
