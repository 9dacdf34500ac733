"""N>1 path on CPU: two gloo ranks shard the tile list, and the job throughput uses the slowest rank."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snap_b200 import parallel
    tiles = parallel.shard_tiles(list(range(7)), rank, world)
    secs = 1.0 if rank == 0 else 2.5            # rank 1 is the slow one
    parallel.barrier()
    thr = parallel.job_throughput(len(tiles), secs)
    mx = parallel.max_over_ranks(secs)
    out.put((rank, tiles, thr, mx))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_gloo_sharding_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, t0, thr0, mx0), (r1, t1, thr1, mx1) = res
    assert t0 == [0, 1, 2] and t1 == [3, 4, 5, 6]        # disjoint, complete, balanced
    assert mx0 == mx1 == 2.5
    assert abs(thr0 - 7 / 2.5) < 1e-9 and thr0 == thr1   # whole-job units / slowest rank


def test_single_process_fallbacks():
    from snap_b200 import parallel
    assert parallel.shard_tiles(list(range(5)), 0, 1) == [0, 1, 2, 3, 4]
    assert parallel.max_over_ranks(3.0) == 3.0 and parallel.job_throughput(4, 2.0) == 2.0
