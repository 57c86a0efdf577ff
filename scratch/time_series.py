"""BASELINE configs[4] at full size: a time series of EW-sized pairs (10400 x 10400, 300 x 300 grid, 3 angles)
through sharding.use_mcc_series -- pairs dealt round-robin to the ranks, two contexts per rank, host images in
pinned memory, every pair uploaded inside the timed region.  Launch with python (1 GPU) or torchrun (N GPUs).

Timing is END TO END on the host clock (device-synchronised and barrier-bracketed on both sides, max over
ranks): the work runs on several streams of several contexts, so no single stream's events bracket it.
Only `--distinct` different pairs are synthesised (the series cycles through them); a seeded sample of every
distinct pair is compared with the exact C oracle (tests/ and benches may use oracle/)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=16)
    ap.add_argument("--distinct", type=int, default=2)
    ap.add_argument("--side", type=int, default=0)
    ap.add_argument("--grid", type=int, default=0)
    ap.add_argument("--contexts", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=3)
    ap.add_argument("--check", type=int, default=1500)
    ap.add_argument("--gather", default="all", choices=["all", "root"])
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from sea_ice_drift_b200 import synthetic as syn
    from sea_ice_drift_b200.sharding import use_mcc_series, shard_pairs
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from sea_ice_drift_b200.sharding import bind_rank_to_gpu
    bound = None if os.environ.get("SID_NO_BIND") else bind_rank_to_gpu(local)
    distinct = []
    for k in range(args.distinct):
        img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg5", seed=k, side=args.side or None, grid=args.grid or None)
        p1 = torch.from_numpy(img1).pin_memory()
        p2 = torch.from_numpy(img2).pin_memory()
        distinct.append((p1.numpy(), p2.numpy(), c1, r1, c2, r2, b, (p1, p2)))
    pairs = [distinct[k % args.distinct][:7] for k in range(args.pairs)]
    n_total = sum(len(p[2]) for p in pairs)
    angles, s = cfg["angles"], cfg["img_size"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    use_mcc_series(pairs[:2 * world], s, 0.0, n_contexts=args.contexts, angles=angles)      # warm-up (allocations)
    times = []
    for _ in range(args.repeat):
        barrier()
        t0 = time.perf_counter()
        tables = use_mcc_series(pairs, s, 0.0, n_contexts=args.contexts, angles=angles,
                                gather=True if args.gather == "all" else "root")
        barrier()
        times.append(time.perf_counter() - t0)
    # breakdown: the rank's own pairs without the final exchange, then the exchange alone
    from sea_ice_drift_b200.sharding import _gather_series
    parts = []
    for _ in range(args.repeat):
        barrier()
        t0 = time.perf_counter()
        local = use_mcc_series(pairs, s, 0.0, n_contexts=args.contexts, angles=angles, gather=False)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        barrier()
        t2 = time.perf_counter()
        if world > 1:
            _gather_series(local, shard_pairs(len(pairs), world, rank), len(pairs), dist, root_only=args.gather == "root")
        barrier()
        parts.append((t1 - t0, time.perf_counter() - t2))
    bd = torch.tensor([min(p[0] for p in parts), min(p[1] for p in parts)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(bd, op=dist.ReduceOp.MAX)
    bd = bd.cpu().numpy()
    t = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs = float(t.item())

    # parity: every rank checks a seeded sample of the distinct pairs against the exact oracle
    bad = 0
    checked = 0
    max_dr = max_dh = 0.0
    if args.check and rank == 0:
        from oracle import c_oracle as co
        from tests.helpers import classify
        rng = np.random.default_rng(0)
        for k in range(args.distinct):
            img1, img2, c1, r1, c2, r2, b = pairs[k]
            idx = np.sort(rng.choice(len(c1), size=min(args.check, len(c1)), replace=False))
            ref, _ = co.use_mcc_batch(c1[idx], r1[idx], c2[idx], r2[idx], b[idx], img1, img2, s, 0.0, angles=angles)
            res = classify(tables[k][idx], ref)
            bad += len(res["unexplained"]) + (0 if res["nan_equal"] else 1)
            max_dr = max(max_dr, res["max_dr"]); max_dh = max(max_dh, res["max_dh"])
            checked += len(idx)
    if rank == 0:
        print(json.dumps({"workload": "cfg5 time series", "pairs": args.pairs, "distinct_pairs": args.distinct,
                          "points_per_pair": len(pairs[0][2]), "vectors": n_total, "n_gpus": world,
                          "contexts_per_gpu": args.contexts, "gather": args.gather, "seconds": round(secs, 5),
                          "vectors_per_s": round(n_total / secs, 1), "ms_per_pair": round(1e3 * secs / args.pairs, 3),
                          "all_times_s": [round(x, 5) for x in times],
                          "compute_and_copies_s": round(float(bd[0]), 5), "gather_s": round(float(bd[1]), 5), "host_images": "pinned", "cores_bound": len(bound) if bound else 0,
                          "includes": "image H2D, point H2D, result D2H, all-gather of tables",
                          "parity_checked": checked, "parity_unexplained": bad,
                          "parity_max_dr": max_dr, "parity_max_dh_rel": max_dh}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
