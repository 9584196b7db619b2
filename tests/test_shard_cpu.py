"""CPU, world_size 2 over gloo: the only multi-rank logic on the path — contiguous sharding of
independent clips and the max-over-ranks of the timings (no data-path collective)."""
import os
import socket
import sys

import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bench.shard_range(65, rank, world)
    mx = bench.dist_max(1.0 + rank, dist)
    dist.barrier()
    q.put((rank, lo, hi, mx))
    dist.destroy_process_group()


def test_shards_cover_all_clips_once_and_time_is_max_over_ranks():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 33), (33, 65)]
    assert all(r[3] == 2.0 for r in res)


@pytest.mark.parametrize("n,world", [(32, 1), (64, 2), (7, 4), (256, 8), (3, 8)])
def test_shard_range_partitions(n, world):
    sys.path.insert(0, ROOT)
    import bench
    spans = [bench.shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
