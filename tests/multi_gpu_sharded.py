"""torchrun --nproc-per-node 2: the sharded use_mcc_batch (NCCL all-gather) and the pair-sharded time series
(use_mcc_series) must equal the single-GPU tables."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sea_ice_drift_b200 import synthetic as syn, pmlib
from sea_ice_drift_b200.sharding import use_mcc_batch_sharded, use_mcc_batch_split, use_mcc_series
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=3, side=3000, grid=50)
b = np.floor(np.random.default_rng(5).uniform(20, 41, b.size))
table = use_mcc_batch_sharded(c1, r1, c2, r2, b, img1, img2, 35, 0.0, angles=cfg["angles"])
dist.barrier()
if rank == 0:
    dist_ok = True
single = pmlib.use_mcc_batch(c1, r1, c2, r2, b, img1, img2, 35, 0.0, angles=cfg["angles"])
same = np.array_equal(table, single, equal_nan=True)
flag = torch.tensor([1 if same else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded(%d ranks) == single GPU: %s   (%d points, %d NaN)" % (dist.get_world_size(), bool(flag.item()), len(c1), int(np.isnan(single[:, 0]).sum())))
# north_star's split: slab upload + in-place all-gather of the pair, kernel rows into the gathered result buffer
split = use_mcc_batch_split(c1, r1, c2, r2, b, img1, img2, 35, 0.0, angles=cfg["angles"])
same_split = np.array_equal(split, single, equal_nan=True)
flag = torch.tensor([1 if same_split else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("split(%d ranks: slab upload + NVLink all-gather) == single GPU: %s" % (dist.get_world_size(), bool(flag.item())))
same = same and same_split
# time series: 5 ragged pairs dealt round-robin to the ranks, twice (the second call reuses the staging buffers)
items = []
for k in range(5):
    i1, i2, pc1, pr1, pc2, pr2, pb, _ = syn.make_config("cfg2", seed=40 + k, side=700 + 64 * k, grid=8 + 4 * k)
    items.append((i1, i2, pc1, pr1, pc2, pr2, pb))
same2 = True
for _ in range(2):
    tables = use_mcc_series(items, 35, 0.0, angles=cfg["angles"])
    for it, t in zip(items, tables):
        one = pmlib.use_mcc_batch(it[2], it[3], it[4], it[5], it[6], it[0], it[1], 35, 0.0, angles=cfg["angles"])
        same2 = same2 and np.array_equal(t, one, equal_nan=True)
flag = torch.tensor([1 if same2 else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("series(%d ranks) == single GPU: %s   (%d pairs)" % (dist.get_world_size(), bool(flag.item()), len(items)))
dist.destroy_process_group()
sys.exit(0 if same and same2 else 1)
