"""CPU: the N > 1 path.  Streams shard over ranks with no data-path collective; the
only communication is the timing reduction bench.py does.  World size 2 over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from crispy_b200.shard import job_rate, stream_block


def test_blocks_cover_every_stream_once():
    for n, w, g in ((1024, 8, 1), (1000, 8, 1), (7, 8, 1), (8192, 4, 2), (10, 4, 2), (0, 2, 1)):
        blocks = [stream_block(n, w, r, g) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
            assert a1 == b0 and a0 <= a1
        sizes = [b - a for a, b in blocks]
        assert max(sizes) - min(sizes) <= g and all(s % g == 0 for s in sizes)
    with pytest.raises(ValueError):
        stream_block(9, 2, 0, group=2)  # a mic/app pair would be split
    with pytest.raises(ValueError):
        stream_block(8, 2, 2)


def _worker(rank, world, port, n_streams, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, last = stream_block(n_streams, world, rank, group=2)
    # every rank "processes" only its own block; the data never leaves the rank
    mine = torch.arange(first, last, dtype=torch.int64)
    local_sum = torch.tensor([int(mine.sum()), last - first], dtype=torch.int64)
    t = torch.tensor([0.25 * (rank + 1)], dtype=torch.float64)  # pretend timings; the job time is the max
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(local_sum, op=dist.ReduceOp.SUM)  # test-only check that the blocks tile the job
    if rank == 0:
        out_q.put((float(t.item()), int(local_sum[0]), int(local_sum[1])))
    dist.destroy_process_group()


def test_two_ranks_partition_and_reduce_like_bench():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 1026
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t_max, id_sum, count = q.get(timeout=10)
    assert t_max == 0.5 and count == n and id_sum == n * (n - 1) // 2
    assert job_rate([513 * 60.0, 513 * 60.0], [0.25, 0.5]) == n * 60.0 / 0.5


def test_cpulist_parser_and_numa_binding_is_harmless():
    from crispy_b200.shard import bind_to_gpu_numa_node, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == []
    import os
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa_node("0000:ff:1f.7") == []  # no such device: affinity untouched
    assert os.sched_getaffinity(0) == before

