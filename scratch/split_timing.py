"""torchrun: strong-scaling split of one cfg2 pair, per-call wall time (max over ranks) + check against rank 0's own table."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from sea_ice_drift_b200 import synthetic as syn, sharding, _lib
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sharding.bind_rank_to_gpu(local)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(name, seed=0)
img1 = torch.from_numpy(img1).pin_memory().numpy(); img2 = torch.from_numpy(img2).pin_memory().numpy()
s, angles = cfg["img_size"], cfg["angles"]
for _ in range(3):
    table = sharding.use_mcc_batch_split(c1, r1, c2, r2, b, img1, img2, s, 0.0, angles=angles, device=local)
dist.barrier(); torch.cuda.synchronize()
ts = []
for _ in range(10):
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    table = sharding.use_mcc_batch_split(c1, r1, c2, r2, b, img1, img2, s, 0.0, angles=angles, device=local)
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ts.append(float(t.item()))
# component timing on the stream (rank 0's view)
ctx = _lib.default_context(local)
stream = torch.cuda.current_stream()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
ctx.set_stream(stream.cuda_stream)
dist.barrier(); torch.cuda.synchronize()
ev[0].record(stream)
sharding.replicate_pair(img1, img2, dist, ctx)
ev[1].record(stream)
ctx.set_stream(None)
torch.cuda.synchronize()
if rank == 0:
    one = _lib.Context(local)
    one.set_pair(img1, img2)
    ref = one.run(c1, r1, c2, r2, b, s, angles, 0.0)
    print("%s split over %d GPUs: %.3f ms per pair (median of 10, max over ranks; min %.3f); replicate_pair %.3f ms on the stream; equals single GPU: %s"
          % (name, world, 1e3 * np.median(ts), 1e3 * min(ts), ev[0].elapsed_time(ev[1]), np.array_equal(table, ref, equal_nan=True)), flush=True)
dist.barrier()
dist.destroy_process_group()
