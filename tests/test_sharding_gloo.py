"""Multi-rank path on CPU: world_size 2, gloo backend, the per-shard GPU batch
replaced by the CPU oracle.  Checks that the all-gathered table equals the
single-process table and that shards are disjoint, complete and balanced."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_indices_partition_and_balance():
    from sea_ice_drift_b200.sharding import shard_indices
    rng = np.random.default_rng(0)
    border = np.floor(rng.uniform(20, 51, 1001))
    for world in (2, 4, 8):
        parts = [shard_indices(border, world, r) for r in range(world)]
        allidx = np.sort(np.concatenate(parts))
        assert np.array_equal(allidx, np.arange(border.size))
        work = [((2 * border[p] + 1) ** 2).sum() for p in parts]
        assert max(work) / min(work) < 1.02
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from sea_ice_drift_b200 import synthetic as syn
    from sea_ice_drift_b200.sharding import use_mcc_batch_sharded
    from tests.test_host_api import oracle_compute
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=5, side=700, grid=9)
        b = np.floor(np.random.default_rng(1).uniform(20, 31, b.size))
        calls = []

        def compute(*a, **k):
            calls.append(len(a[0]))
            return oracle_compute(*a, **k)
        table = use_mcc_batch_sharded(c1, r1, c2, r2, b, img1, img2, 35, 0.0, compute=compute, angles=[-3, 0, 3])
        q.put((rank, table, calls))
    finally:
        dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    import torch.multiprocessing as mp
    from sea_ice_drift_b200 import synthetic as syn
    from tests.test_host_api import oracle_compute
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=5, side=700, grid=9)
    b = np.floor(np.random.default_rng(1).uniform(20, 31, b.size))
    single = oracle_compute(c1, r1, c2, r2, b, img1, img2, 35, 0.0, angles=[-3, 0, 3])
    for rank, table, calls in got:
        assert np.array_equal(table, single, equal_nan=True)
        assert len(calls) == 1 and abs(calls[0] - len(c1) / 2) <= 1


def _series_pairs():
    """Five small pairs with different grid sizes (ragged tables), the last one lazily loaded."""
    from sea_ice_drift_b200 import synthetic as syn
    items = []
    for k in range(5):
        img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=20 + k, side=500 + 32 * k, grid=5 + k)
        items.append((img1, img2, c1, r1, c2, r2, b))
    last = items[-1]
    items[-1] = lambda: last
    return items


def _series_compute(slot, img1, img2, c1, r1, c2, r2, b):
    from tests.test_host_api import oracle_compute
    return oracle_compute(c1, r1, c2, r2, b, img1, img2, 35, 0.0, angles=[-3, 0, 3])


def _series_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from sea_ice_drift_b200.sharding import use_mcc_series
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        seen = []

        def compute(slot, *a):
            seen.append((slot, len(a[2])))
            return _series_compute(slot, *a)
        tables = use_mcc_series(_series_pairs(), 35, 0.0, n_contexts=2, compute=compute)
        rooted = use_mcc_series(_series_pairs(), 35, 0.0, compute=_series_compute, gather='root')
        assert all(t is not None for t in rooted) if rank == 0 else [t is None for t in rooted] == [True, False] * 2 + [True]
        if rank == 0:
            assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(rooted, tables))
        q.put((rank, tables, seen))
    finally:
        dist.destroy_process_group()


def test_shard_pairs_partition():
    from sea_ice_drift_b200.sharding import shard_pairs
    for world in (1, 2, 3, 8):
        parts = [shard_pairs(16, world, r) for r in range(world)]
        assert sorted(sum(parts, [])) == list(range(16))
        assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert shard_pairs(3, 8, 5) == []


def test_series_single_process_and_error_propagation():
    from sea_ice_drift_b200.sharding import use_mcc_series
    pairs = _series_pairs()
    tables = use_mcc_series(pairs, 35, 0.0, n_contexts=3, compute=_series_compute)
    for item, t in zip(pairs, tables):
        item = item() if callable(item) else item
        assert np.array_equal(t, _series_compute(0, *item), equal_nan=True)
    assert use_mcc_series([], 35, compute=_series_compute) == []

    def boom(slot, *a):
        raise RuntimeError("pair failed")
    with pytest.raises(RuntimeError, match="pair failed"):
        use_mcc_series(pairs, 35, compute=boom)


def test_two_rank_series_gather_equals_single_process():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_series_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    pairs = _series_pairs()
    want = [_series_compute(0, *(it() if callable(it) else it)) for it in pairs]
    for rank, tables, seen in got:
        assert len(tables) == len(want)
        for t, w in zip(tables, want):
            assert np.array_equal(t, w, equal_nan=True)
        assert len(seen) == (3 if rank == 0 else 2)          # pairs 0,2,4 / 1,3 -- nothing computed twice


def test_bind_rank_to_gpu_is_a_no_op_without_a_gpu():
    import os
    from sea_ice_drift_b200.sharding import bind_rank_to_gpu
    before = os.sched_getaffinity(0)
    cores = bind_rank_to_gpu(0)
    assert cores is None or set(cores) <= before          # never widens the affinity, never raises
    if cores is None:
        assert os.sched_getaffinity(0) == before


def _replicate_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from sea_ice_drift_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(3)
        img1 = rng.integers(1, 256, (301, 257), dtype=np.uint8)      # odd sizes: the last slab is padded
        img2 = rng.integers(1, 256, (299, 263), dtype=np.uint8)

        class FakeCtx(object):                                       # no GPU here: only the buffers are checked
            device = 0
        uploaded, bufs = sharding.replicate_pair(img1, img2, dist, ctx=FakeCtx())
        out = []
        for (buf, plan), img in zip(bufs, (img1, img2)):
            full = buf[:plan.rows * plan.pitch].numpy().reshape(plan.rows, plan.pitch)
            out.append((np.array_equal(full[:, :plan.cols], img), int(full[:, plan.cols:].max()), plan.row_range(rank)))
        q.put((rank, uploaded, out))
    finally:
        dist.destroy_process_group()


def test_replicate_pair_slabs_fill_every_rank():
    """north_star's replication: each rank contributes 1/world of the rows, the all-gather completes the padded
    image on every rank (gloo stand-in for the NCCL all-gather over NVLink)."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_replicate_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = 0
    for rank, uploaded, out in got:
        total += uploaded
        for same, pad_max, (r0, r1) in out:
            assert same and pad_max == 0 and r1 > r0
    assert total == 301 * 257 + 299 * 263                            # every byte crossed "PCIe" exactly once


def test_split_plan_geometry():
    from sea_ice_drift_b200.sharding import SplitPlan
    for rows, world in ((10400, 8), (301, 2), (7, 8)):
        plan = SplitPlan(rows, 10400, world)
        assert plan.pitch % 16 == 0 and plan.pitch >= 10400 + 16
        ranges = [plan.row_range(r) for r in range(world)]
        assert ranges[0][0] == 0 and max(r1 for _, r1 in ranges) == rows
        assert all(a[1] == b[0] or b[0] == rows for a, b in zip(ranges, ranges[1:]))
        assert plan.gather_bytes == world * plan.slab and plan.alloc_bytes >= plan.bytes
