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


def parse_cpulist(text: str) -> list[int]:
    """"0-3,8,10-11" -> [0, 1, 2, 3, 8, 10, 11] (the format of /sys/.../local_cpulist)."""
    cpus: list[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(pci_bus_id: str) -> list[int]:
    """Pin the calling process to the CPUs next to the GPU `pci_bus_id` ("0000:1b:00.0").  With one process per GPU
    the pinned staging buffers of process_streams_host are then first-touched on the GPU's own NUMA node, so eight
    ranks do not all stream through one socket's memory controllers.  Returns the CPU list ([] = left alone)."""
    import os
    try:
        bdf = pci_bus_id.lower()
        if bdf.count(":") == 1:
            bdf = "0000:" + bdf
        if len(bdf.split(":")[0]) == 8:  # nvml-style 00000000:1B:00.0
            bdf = bdf[4:]
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            cpus = parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return []

