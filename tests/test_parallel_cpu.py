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
    # gradient mean over ranks in flat buckets (row 21): rank r holds (r+1) * base
    g = torch.Generator().manual_seed(3)
    base = {"bev_mapper": {"fusion": {"kernel": torch.randn(257, 16, generator=g), "bias": torch.randn(16, generator=g)}},
            "head": [torch.randn(5, generator=g).to(torch.bfloat16), torch.randn(3, 3, generator=g)]}
    scale = lambda t: {k: scale(v) for k, v in t.items()} if isinstance(t, dict) else (
        [scale(v) for v in t] if isinstance(t, list) else t * (rank + 1))
    grads = scale(base)
    calls_one = parallel.pmean_tree(grads)
    want = lambda t: {k: want(v) for k, v in t.items()} if isinstance(t, dict) else (
        [want(v) for v in t] if isinstance(t, list) else t.float() * 1.5)
    flat = lambda t: torch.cat([x.float().reshape(-1) for _, x in parallel._leaves(t)])
    err = float((flat(grads) - flat(want(base))).abs().max())
    grads2 = scale(base)
    calls_small = parallel.pmean_tree(grads2, bucket_bytes=64)
    err2 = float((flat(grads2) - flat(grads)).abs().max())
    out.put((rank, tiles, thr, mx, calls_one, err, calls_small, err2))
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
    (r0, t0, thr0, mx0, *g0), (r1, t1, thr1, mx1, *g1) = res
    assert g0 == g1
    calls_one, err, calls_small, err2 = g0
    assert calls_one == 1 and err < 2e-2                   # one bucket; bf16 leaf rounds on write-back
    assert calls_small == 3 and err2 == 0.0                # bucketing does not change the result
    assert t0 == [0, 1, 2] and t1 == [3, 4, 5, 6]        # disjoint, complete, balanced
    assert mx0 == mx1 == 2.5
    assert abs(thr0 - 7 / 2.5) < 1e-9 and thr0 == thr1   # whole-job units / slowest rank


def test_single_process_fallbacks():
    from snap_b200 import parallel
    assert parallel.shard_tiles(list(range(5)), 0, 1) == [0, 1, 2, 3, 4]
    assert parallel.pmean_tree({"a": torch.ones(3)}) == 0
    assert parallel.max_over_ranks(3.0) == 3.0 and parallel.job_throughput(4, 2.0) == 2.0


def _grad_mean_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    from snap_b200 import parallel
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the gradient tree of the MLP semantic head (Dense_0..2), different on every rank
    g = torch.Generator().manual_seed(100 + rank)
    tree = {f"Dense_{i}": {"kernel": torch.randn((128, 128 if i < 2 else 32), generator=g), "bias": torch.randn((128 if i < 2 else 32,), generator=g)}
            for i in range(3)}
    calls = parallel.pmean_tree(tree)
    q.put((rank, calls, {k: {n: t.numpy().copy() for n, t in v.items()} for k, v in tree.items()}))   # NumPy: picklable by value
    dist.destroy_process_group()


def test_head_gradient_mean_world_size_2():
    """trainer.py:231-234 (pmean of the head gradients) on two gloo ranks: one bucketed all-reduce, identical means."""
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_mean_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == res[1][1] == 1
    for k in res[0][2]:
        for n in res[0][2][k]:
            assert (res[0][2][k][n] == res[1][2][k][n]).all()
    g0, g1 = torch.Generator().manual_seed(100), torch.Generator().manual_seed(101)
    a, b = torch.randn((128, 128), generator=g0), torch.randn((128, 128), generator=g1)
    assert torch.allclose(torch.from_numpy(res[0][2]["Dense_0"]["kernel"]), (a + b) / 2, atol=1e-6)


def _dp_case():
    """Deterministic inputs of the DP-equivalence test: 4 scenes, MLP head, labels."""
    import numpy as np
    rng = np.random.default_rng(0)
    B, G = 4, 8
    p = {f"Dense_{i}": {"kernel": (rng.standard_normal((16, 16 if i < 2 else 12)) * 0.3).astype(np.float32),
                        "bias": (rng.standard_normal(16 if i < 2 else 12) * 0.1).astype(np.float32)} for i in range(3)}
    feats = rng.standard_normal((B, G, G, 16)).astype(np.float32)
    valid = rng.random((B, G, G)) < 0.8
    la, le = rng.integers(0, 5, (B, G, G)), rng.integers(0, 4, (B, G, G))
    va = rng.random((B, G, G)) < 0.9
    mi = rng.random((B, G, G, 3)) < 0.2
    return p, feats, valid, la, va, le, mi


def _dp_grads(sl):
    import numpy as np
    import torch
    from oracle import semantic_net as osn
    p, feats, valid, la, va, le, mi = _dp_case()
    tp = {k: {n: torch.from_numpy(v[n]).requires_grad_(True) for n in v} for k, v in p.items()}
    logits = osn.mlp_head_forward_torch(torch.from_numpy(feats[sl]), valid[sl], tp)
    loss, _ = osn.total_loss_torch(logits, la[sl], va[sl], le[sl], mi[sl], valid[sl], 5, 4)
    loss.backward()
    return {k: {n: t.grad.clone() for n, t in v.items()} for k, v in tp.items()}


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snap_b200 import parallel
    per = 4 // world
    grads = _dp_grads(slice(rank * per, (rank + 1) * per))     # this rank's scenes (trainer.py:452-464 shards the batch)
    parallel.pmean_tree(grads)                                 # trainer.py:231-234
    q.put((rank, {k: {n: t.numpy().copy() for n, t in v.items()} for k, v in grads.items()}))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_data_parallel_step_equivalence_world_size_2():
    """the mean over ranks of the per-shard gradients (2 ranks x 2 scenes) equals the gradient of the whole
    batch (1 x 4 scenes): the loss is a mean of per-example losses (trainer.py:221) and shards are equally sized."""
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=150) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    full = _dp_grads(slice(0, 4))
    for k in full:
        for n in full[k]:
            ref = full[k][n].numpy()
            for r in range(2):
                assert np.abs(res[r][1][k][n] - ref).max() <= 1e-6 * (1 + np.abs(ref).max()), (k, n)


def _bucket_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from snap_b200 import parallel
    shapes = [(257, 16), (16,), (3, 3, 5), (1,)]
    b = parallel.GradBucket(shapes, "cpu")
    g = torch.Generator().manual_seed(11)
    base = [torch.randn(s, generator=g) for s in shapes]
    for v, t in zip(b.views, base):          # "backward kernels" write straight into the bucket's views
        v.copy_(t * (rank + 1))
    ptr0 = b.flat.data_ptr()
    calls = b.allreduce_mean()
    err = max(float((v - t * 1.5).abs().max()) for v, t in zip(b.views, base))
    aligned = all(v.data_ptr() % 16 == 0 for v in b.views)
    finite = bool(b.all_finite())
    if rank == 1:
        b.views[2].view(-1)[7] = float("nan")
    b.allreduce_mean()                       # a NaN on one rank reaches every rank through the mean
    out.put((rank, calls, err, aligned, finite, bool(b.all_finite()), b.flat.data_ptr() == ptr0, b.nbytes))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_grad_bucket_single_in_place_all_reduce():
    """`parallel.GradBucket`: gradients live in one flat fp32 buffer, the mean over ranks (trainer.py:231-234) is ONE in-place
    collective with no allocation, and the non-finite guard (trainer.py:260-276) sees a NaN produced on any rank."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=90) for _ in range(2))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, calls, err, aligned, finite, finite_after, same_buffer, nbytes in res:
        assert calls == 1 and err < 1e-6 and aligned and finite and not finite_after and same_buffer
        assert nbytes == 4 * (4112 + 16 + 48 + 4)       # every leaf padded to 16 bytes


def test_grad_bucket_without_process_group_is_identity():
    from snap_b200 import parallel
    b = parallel.GradBucket([(4, 4), (3,)], "cpu")
    b.views[0].fill_(2.0)
    assert b.allreduce_mean() == 0 and float(b.views[0].sum()) == 32.0 and bool(b.all_finite())
