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
