"""N>1 host logic on CPU: world_size-2 gloo.  The path shards image-wise with NO data-path collective; the
only cross-rank operations are the barrier and the max-over-ranks of the timed region, which is what is tested here."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = bench.shard_range(512, rank, world)          # C5: contiguous shards [g*512/n, (g+1)*512/n)
    my_time = 0.25 * (rank + 1)                           # rank 1 is the slow one
    t = bench.dist_max(my_time)
    frames = bench.dist_sum(hi - lo)
    dist.barrier()
    if rank == 0:
        line = bench.result_line(frames=(hi - lo), seconds=t, n_gpus=world, steps=1, warmup=3, extra={})
        q.put((t, frames, line["value"], lo, hi))
    else:
        q.put((t, frames, None, lo, hi))
    dist.destroy_process_group()


def test_two_rank_sharding_and_max_time():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spans = sorted((r[3], r[4]) for r in res)
    assert spans == [(0, 256), (256, 512)]
    for t, frames, value, lo, hi in res:
        assert t == pytest.approx(0.5)        # max over ranks, not this rank's own time
        assert frames == 512                  # every frame is owned by exactly one rank
        if value is not None:
            assert value == pytest.approx(256 * 2 / 0.5)
