"""Partition of independent streams over the GPUs of one box.

One `DenoiseState` per stream and no shared mutable state (src-tauri/src/audio.rs:203), so the
partition is a contiguous block of streams per rank and the data path needs no collective.  Sources
of one meeting (mic + app audio, commands/recording.rs:260-264) are kept on one rank so the dual-mono
mix never crosses GPUs.  torch.distributed is used by callers only for the timing barrier and the
max-over-ranks reduction of the measured time.
"""
from __future__ import annotations


def stream_block(n_streams: int, world_size: int, rank: int, group: int = 1) -> tuple[int, int]:
    """[first, last) of the streams rank `rank` owns.  `group` streams that belong together (2 for a
    mic/app pair) are never split across ranks; blocks differ in size by at most one group."""
    if n_streams < 0 or world_size < 1 or not (0 <= rank < world_size) or group < 1 or n_streams % group:
        raise ValueError("bad partition arguments")
    units = n_streams // group
    base, extra = divmod(units, world_size)
    first = rank * base + min(rank, extra)
    count = base + (1 if rank < extra else 0)
    return first * group, (first + count) * group


def job_rate(units_per_rank: list[float], seconds_per_rank: list[float]) -> float:
    """whole-job throughput: all units processed / the slowest rank's time (bench.py contract)."""
    return sum(units_per_rank) / max(seconds_per_rank)
