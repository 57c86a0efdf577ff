#!/usr/bin/env python
"""Benchmark of the pattern-matching (MCC) hot path -- BASELINE.json's metric
"PM grid vectors/sec" on configs[1] (Sentinel-1 EW-sized synthetic pair 10400 x 10400,
rotational drift, 200 x 200 PM grid, img_size 35, angles [-3, 0, 3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over the whole PM grid of one image pair.
  value   : whole-job vectors/s with the pair and the point arrays already resident in HBM
            (device-timed with CUDA events on the launching stream, max over ranks).
  e2e     : the same metric through the C-ABI call a user makes (sid_run_pair) with
            HOST buffers: pinned-host -> device copy of the image pair and the point arrays and
            the device -> host read of the result table are inside the timed region.
  roofline: algorithmic FLOPs (sum over points and angles of 2 s^2 R^2) per launch
            over the kernel's average duration, against the measured dense tensor peak of
            MEASURED_PEAKS.json (the multiply-adds run as tcgen05 kind::i8 MMAs); roofline_fma repeats it
            against the FP32-FMA peak that BASELINE.json's metric names.
  configs : the other BASELINE configurations at full size: cfg1 (borders of a real ORB first guess), cfg3, cfg4
            device-resident; cfg5 (time series) end to end through sharding.use_mcc_series, this GPU's share of the pairs.
  drop_in : the whole call a user of the reference makes -- pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2) on
            configs[0] with NumPy images -- wall clock, next to the UNMODIFIED reference's pattern_matching on the same
            inputs (its own ORB matches; timed in a child process before CUDA is initialised).
  cpu_baseline / --impl reference: the UNMODIFIED reference's own per-point loop (pmlib.use_mcc_mp through a
            fork Pool over all host cores, from the oracle/_ref copy placed by oracle/build_ref.py; kind
            "reference") on a bounded sample of the same workload -- the NumPy/cv2/scipy port
            (oracle/pm_oracle.py, kind "port") only where that copy is absent.
  parity  : the GPU table against the exact CPU oracle AND against the reference on seeded samples.
Multi-GPU (torchrun, one rank per GPU): `value` / `e2e` are weak scaling -- every rank matches the full grid of
its own image pair (the time-series case, BASELINE configs[4]); no data-path collective.  The `strong` block is
north_star's split of ONE pair: every rank uploads 1/N of the rows, an in-place NCCL all-gather over NVLink
completes the pair on every GPU, the points are dealt over the ranks and one all-gather returns the table.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PM grid vectors/sec"
UNIT = "vectors/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--side", type=int, default=0, help="shrink the image side (debug)")
    ap.add_argument("--grid", type=int, default=0, help="shrink the grid side (debug)")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="points timed for cpu_baseline")
    ap.add_argument("--ref-sample", type=int, default=4000, help="points per step of --impl reference")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the cfg1 / cfg3 / cfg4 block")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling block at N > 1")
    ap.add_argument("--drop-in-reference", default="", metavar="NPZ",
                    help="(internal, child process) time the reference's own pattern_matching on configs[0] and save the matches")
    return ap.parse_args()


def workload_description(name, cfg, n):
    return ("%s: synthetic speckle pair %dx%d uint8, known %s drift, %d-point PM grid (%dx%d requested), "
            "img_size=%d, border=%s, angles=%s" % (name, cfg["side"], cfg["side"], cfg["warp"][0], n, cfg["grid"],
                                                   cfg["grid"], cfg["img_size"], cfg["border"], cfg["angles"]))


def flops_per_point(img_size, border, n_angles):
    s = img_size
    w = 2 * (s // 2) + 2 * np.asarray(border, dtype=np.float64) + 1
    r = w - s + 1
    return n_angles * 2.0 * s * s * r * r


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.lines, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.kill()
        self.proc.wait()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [l for (t, l) in self.lines if t0 - 0.05 <= t <= t1 + 0.15] or [l for (_, l) in self.lines]
        for l in rows:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_kind():
    """"reference": the unmodified reference (oracle/_ref copy or /root/reference) can be loaded; else "port"."""
    from oracle import ref_runner
    return "reference" if ref_runner.available() else "port"


def cpu_port_rate(pts, img1, img2, img_size, angles, sample, threads, seed=0):
    """vectors/s of the reference's CPU loop on `sample` points of the workload: the reference's own use_mcc_mp
    when the reference copy is present, else the port that calls the same third-party routines."""
    n = len(pts[0])
    sel = np.sort(np.random.default_rng(seed).choice(n, min(sample, n), replace=False))
    sub = [p[sel] for p in pts]
    if cpu_kind() == "reference":
        from oracle import ref_runner
        t0 = time.perf_counter()
        rows = ref_runner.run_reference_points(*sub, img1, img2, img_size, 0.0, threads=threads, angles=angles)
    else:
        from oracle import pm_oracle
        t0 = time.perf_counter()
        rows = pm_oracle.run_points(*sub, img1, img2, img_size, 0.0, threads=threads, angles=angles)
    dt = time.perf_counter() - t0
    return len(sel) / dt, len(sel), dt, rows, sel


DROP_IN = dict(side=2000, grid=50, img_size=35, angles=[0])


def drop_in_scene(syn):
    """BASELINE configs[0] for the whole drop-in call: the pair, its affine domains and the 50 x 50 lon / lat grid."""
    img1, img2, _, _, _, _, _, _ = syn.make_config("cfg1", seed=0, side=DROP_IN["side"])
    n1, n2 = syn.ArrayDomain(img1), syn.ArrayDomain(img2)
    gx, gy = np.meshgrid(np.linspace(100, DROP_IN["side"] - 100, DROP_IN["grid"]), np.linspace(100, DROP_IN["side"] - 100, DROP_IN["grid"]))
    lon, lat = n2.transform_points(gx, gy)
    return img1, img2, n1, n2, lon, lat


def drop_in_reference_child(path):
    """Child process (no CUDA): the UNMODIFIED reference end to end on configs[0] -- its own ORB feature tracking
    (find_key_points / get_match_coords / lstsq_filter), then its pattern_matching (prepare_first_guess + fork Pool over
    use_mcc_mp + post-processing) on all host cores.  Saves the matches for the product's run; prints one JSON line."""
    import contextlib
    import io
    from sea_ice_drift_b200 import synthetic as syn
    from oracle import ref_runner
    img1, img2, n1, n2, lon, lat = drop_in_scene(syn)
    x1, y1, x2, y2 = ref_runner.orb_first_guess(img1, img2)
    np.savez(path, x1=x1, y1=y1, x2=x2, y2=y2)
    pm = ref_runner.reference_module()
    import nansat
    nansat.NSR = lambda srs=None: srs
    pm.NSR = nansat.NSR
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res = pm.pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2, threads=threads, angles=DROP_IN["angles"], img_size=DROP_IN["img_size"])
    dt = time.perf_counter() - t0
    print(json.dumps({"seconds": dt, "threads": threads, "valid": int(np.isfinite(res[0]).sum()), "matches": int(len(x1)),
                      "checksum": float(np.nansum(np.abs(res[0])) + np.nansum(np.abs(res[1])) + np.nansum(res[2]))}), flush=True)
    return 0


def drop_in_block(syn, ref_line):
    """The call a user of the reference makes -- pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2) on configs[0] with
    plain NumPy images -- timed end to end (first guess on the device, upload, matching, post-processing epilogue),
    next to the unmodified reference on the same inputs (timed by the child process before CUDA was initialised)."""
    import contextlib
    import io
    from sea_ice_drift_b200 import pmlib
    m = np.load(ref_line["npz"])
    img1, img2, n1, n2, lon, lat = drop_in_scene(syn)
    kw = dict(angles=DROP_IN["angles"], img_size=DROP_IN["img_size"])

    def checksum(res):      # u, v (displacements) and the winning angles summed: equal to 1e-9 relative means the same vectors
        return float(np.nansum(np.abs(res[0])) + np.nansum(np.abs(res[1])) + np.nansum(res[2]))

    def timed(**extra):
        ts, res = [], None
        for _ in range(6):
            t0 = time.perf_counter()
            with contextlib.redirect_stdout(io.StringIO()):
                res = pmlib.pattern_matching(lon, lat, n1, m["x1"], m["y1"], n2, m["x2"], m["y2"], **kw, **extra)
            ts.append(time.perf_counter() - t0)
        same = bool(int(np.isfinite(res[0]).sum()) == ref_line["valid"] and
                    abs(checksum(res) - ref_line["checksum"]) <= 1e-9 * max(1.0, abs(ref_line["checksum"])))
        return float(np.median(ts[1:])), ts[0], int(np.isfinite(res[0]).sum()), same

    mine, first, valid, same = timed()
    dev, _, valid_dev, same_dev = timed(first_guess="device")
    return {"call": "pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2, angles=[0], img_size=35) on configs[0]: 2000 x 2000 pair, "
                    "%d ORB matches (the reference's own feature tracking), 50 x 50 grid; NumPy images in, seven (50, 50) grids out"
                    % ref_line["matches"],
            "ms": mine * 1e3, "first_call_ms": first * 1e3, "vectors_per_s": valid / mine, "valid": valid,
            "reference_s": ref_line["seconds"], "reference_threads": ref_line["threads"], "reference_valid": ref_line["valid"],
            "speedup": ref_line["seconds"] / mine, "same_vectors": same,
            "ms_first_guess_device": dev * 1e3, "speedup_first_guess_device": ref_line["seconds"] / dev,
            "same_vectors_first_guess_device": same_dev,
            "note": "default first_guess='auto': ORB keypoints of pyramid level 0 are integer pixels, so under this scene's identity "
                    "geolocation some Delaunay cells are cocircular and the triangulation is not unique; 'auto' detects that and takes "
                    "the SciPy / Qhull path (~85 ms of the call) so that the first guess is the reference's in every case. Keypoints in "
                    "generic position (any real geolocation) and first_guess='device' use the triangulation-free device interpolant"}


def drop_in_ew_block(syn, img1, img2, angles, n_keypoints=50000, grid=200):
    """The same call at the headline size: pattern_matching on the resident EW-sized pair (plain NumPy images, so the call
    uploads them), 50 000 synthetic feature-tracking matches in generic (non-integer) position that follow the known drift
    with 0.8 px noise, 200 x 200 grid, angles [-3, 0, 3].  The reference is not run here (prepare_first_guess alone takes
    ~24 s at this size)."""
    import contextlib
    import io
    from sea_ice_drift_b200 import pmlib
    rng = np.random.default_rng(7)
    side = img1.shape[0]
    m = syn.rotation_matrix(img1.shape, syn.CONFIGS["cfg2"]["warp"][1])
    kx, ky = rng.uniform(40, side - 40, n_keypoints), rng.uniform(40, side - 40, n_keypoints)
    k2x, k2y = syn.apply_affine(m, kx, ky)
    k2x, k2y = k2x + rng.normal(0, 0.8, n_keypoints), k2y + rng.normal(0, 0.8, n_keypoints)
    n1, n2 = syn.ArrayDomain(img1), syn.ArrayDomain(img2)
    gx, gy = np.meshgrid(np.linspace(150, side - 150, grid), np.linspace(150, side - 150, grid))
    lon, lat = n2.transform_points(gx, gy)
    ts, res = [], None
    for _ in range(5):
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()):
            res = pmlib.pattern_matching(lon, lat, n1, kx, ky, n2, k2x, k2y, angles=angles, img_size=35)
        ts.append(time.perf_counter() - t0)
    mine = float(np.median(ts[1:]))
    valid = int(np.isfinite(res[0]).sum())
    # known answer: the drift of the synthetic pair at the grid points (minus the reference's -1 px template-centre bias)
    ok = np.isfinite(res[0])
    c2, r2 = n2.transform_points(res[5][ok], res[6][ok], 1)
    c1g, r1g = n1.transform_points(lon[ok], lat[ok], 1)
    tx, ty = syn.apply_affine(m, c1g, r1g)
    err = float(np.median(np.hypot(c2 - (tx - 1), r2 - (ty - 1))))
    return {"call": "pattern_matching(lon, lat, n1, x1, y1, n2, x2, y2, angles=[-3, 0, 3], img_size=35): %d x %d NumPy pair, %d matches "
                    "in generic position, %d x %d grid" % (side, side, n_keypoints, grid, grid),
            "ms": mine * 1e3, "first_call_ms": ts[0] * 1e3, "valid": valid, "vectors_per_s": valid / mine,
            "median_error_px_vs_known_drift": err,
            "reference": "not run (prepare_first_guess alone: 23.8 s at this size, BASELINE.md section 2; the per-point loop: cpu_baseline)"}


def main_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (see module doc)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from sea_ice_drift_b200 import synthetic as syn
    img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(args.workload, seed=0, side=args.side or None,
                                                         grid=args.grid or None)
    pts = [c1, r1, c2, r2, b]
    threads = os.cpu_count() or 1
    for w in range(args.warmup):
        cpu_port_rate(pts, img1, img2, cfg["img_size"], cfg["angles"], min(args.ref_sample, 500), threads, seed=100 + w)
    t0 = time.perf_counter()
    done = 0
    for k in range(args.steps):
        _, m, _, _, _ = cpu_port_rate(pts, img1, img2, cfg["img_size"], cfg["angles"], args.ref_sample, threads, seed=k)
        done += m
    dt = time.perf_counter() - t0
    value = done / dt
    sample = "%d seeded random grid points per step (of %d), %d steps" % (min(args.ref_sample, len(c1)), len(c1), args.steps)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload_description(args.workload, cfg, len(c1))},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": cpu_kind(), "sample": sample,
                             "source": "oracle/_ref: unmodified sea_ice_drift/pmlib.py use_mcc_mp through a fork Pool"
                                       if cpu_kind() == "reference" else "oracle/pm_oracle.py (port)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def time_resident(ctx, torch, stream, dev, pts, s, angles, steps, warmup):
    """Device-timed steps of the hot path on the resident pair; returns (ms per step, dominant-kernel ms, table, status)."""
    n = len(pts[0])
    d_pts = torch.from_numpy(np.stack(pts)).to(dev)
    d_out = torch.empty((n, 5), dtype=torch.float64, device=dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    ptrs = [d_pts[k].data_ptr() for k in range(5)]
    max_border = int(np.max(pts[4]))

    def step():
        ctx.run_device(n, *ptrs, max_border, s, angles, 0.0, d_out.data_ptr(), d_status.data_ptr())
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    k = []
    for _ in range(3):
        step()
        k.append(ctx.last_kernel_ms)
    return ms, float(np.mean(k)), d_out.cpu().numpy(), d_status.cpu().numpy()


def reference_parity(out_rows, ref_rows, pts, img1, img2, s, angles):
    """GPU rows against the reference's rows on the same points: exact / tie-explained / unexplained, max |dr|,
    max |dh| and the number of points above north_star's 1e-4 absolute bound in h (tests/helpers.py explains it)."""
    from oracle import c_oracle
    from tests.helpers import classify, make_exact_lookup
    opts = dict(rot_order=0, hes_norm=True, hes_smth=False, mcc_norm=False)
    st = classify(out_rows, ref_rows, make_exact_lookup(c_oracle, pts, img1, img2, s, 0.0, angles, opts))
    return {"points": st["n"], "nan_pattern_equal": bool(st["nan_equal"]), "position_angle_exact": st["exact"],
            "tie_explained": st["ties"], "unexplained": len(st["unexplained"]), "max_abs_dr": st["max_dr"],
            "max_abs_dh": st["max_dh_abs"], "n_dh_above_1e-4": st["n_dh_gt_1e4"],
            "note": "h above 1e-4 absolute only where |h| > 10: cv2's float32 DFT noise (2e-6 in r) amplified by "
                    "(hes - median) / std; the exact-integer GPU side is the more accurate one"}


def main_ours(args):
    import torch
    import torch.distributed as dist
    from sea_ice_drift_b200 import _lib, synthetic as syn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # weak scaling: rank r matches the grid of its own pair (seed r), like one scene of a time series
    img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(args.workload, seed=rank, side=args.side or None,
                                                         grid=args.grid or None)
    n = len(c1)
    s, angles = cfg["img_size"], cfg["angles"]

    # CPU baseline first (rank 0, N = 1 only): its fork Pool must not inherit a live CUDA context
    cpu_run, cpu_extra = None, None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cpu_run = (threads,) + cpu_port_rate([c1, r1, c2, r2, b], img1, img2, s, angles, args.cpu_sample, threads)
        cpu_extra = {"threads_1": cpu_port_rate([c1, r1, c2, r2, b], img1, img2, s, angles, max(200, args.cpu_sample // 40), 1)[0],
                     "threads_5_reference_default": cpu_port_rate([c1, r1, c2, r2, b], img1, img2, s, angles,
                                                                  max(500, args.cpu_sample // 8), 5)[0]}

    # the reference's whole pattern_matching on configs[0], in a child process (its fork Pool and cv2 stay away from CUDA)
    ref_drop_in = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline and not args.no_configs and cpu_kind() == "reference":
        import subprocess
        import tempfile
        npz = os.path.join(tempfile.mkdtemp(prefix="sid_bench_"), "matches.npz")
        try:
            child = subprocess.run([sys.executable, os.path.abspath(__file__), "--drop-in-reference", npz], stdout=subprocess.PIPE,
                                   stderr=subprocess.PIPE, text=True, timeout=600)
            lines = [l for l in child.stdout.splitlines() if l.startswith("{")]
            if child.returncode == 0 and lines:
                ref_drop_in = json.loads(lines[-1])
                ref_drop_in["npz"] = npz
            else:
                sys.stderr.write("drop-in reference child failed: %s\n" % child.stderr[-800:])
        except (OSError, subprocess.SubprocessError, ValueError) as exc:
            sys.stderr.write("drop-in reference child failed: %r\n" % (exc,))

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this benchmark has no CPU path; use --impl reference)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # one process per GPU: keep this rank (and the pinned buffers it allocates from here on) on the GPU's NUMA node
    from sea_ice_drift_b200.sharding import bind_rank_to_gpu
    bound = None if os.environ.get("SID_NO_BIND") else bind_rank_to_gpu(local_rank)
    img1p = torch.from_numpy(img1).pin_memory().numpy()
    img2p = torch.from_numpy(img2).pin_memory().numpy()

    ctx = _lib.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.set_pair(img1p, img2p)
    d_pts = torch.from_numpy(np.stack([c1, r1, c2, r2, b])).to(dev)
    d_out = torch.empty((n, 5), dtype=torch.float64, device=dev)
    d_status = torch.empty(n, dtype=torch.int32, device=dev)
    max_border = int(b.max())
    ptrs = [d_pts[k].data_ptr() for k in range(5)]

    def step_resident():
        ctx.run_device(n, *ptrs, max_border, s, angles, 0.0, d_out.data_ptr(), d_status.data_ptr())

    # ---- device-resident timing ("value")
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier(); torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.25 if rank == 0 else 0.0)
    barrier(); torch.cuda.synchronize()
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    torch.cuda.synchronize()
    t_wall1 = time.time()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - launches0
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    total_points = sum_over_ranks(float(n))
    value = total_points * args.steps / (ms_total * 1e-3)
    ms_per_step = ms_total / args.steps

    out = d_out.cpu().numpy()
    status = d_status.cpu().numpy()
    flops_step = float(flops_per_point(s, b[status == 1], len(angles)).sum())
    # dominant kernel = sid::pm_points_kernel (a step also runs the light pm_tail_kernel): its own device time,
    # CUDA events recorded by the library on the launching stream around that launch, averaged over 5 launches
    k_samples = []
    for _ in range(5):
        step_resident()
        k_samples.append(ctx.last_kernel_ms)
    kernel_ms = float(np.mean(k_samples))
    kernel_name = ctx.last_kernel_name

    # ---- end to end through the C ABI with host buffers ("e2e")
    ctx.set_stream(None)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        host_out = ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0)
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_out = ctx.run_pair(img1p, img2p, c1, r1, c2, r2, b, s, angles, 0.0)
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_points * e2e_steps / dt_e2e
    h2d = int(img1.nbytes + img2.nbytes + n * 5 * 8 + n * 4 + len(angles) * 5 * 8)
    d2h = int(n * 5 * 8)
    same_as_resident = bool(np.array_equal(host_out, out, equal_nan=True))
    # the same call with plain (pageable) NumPy images -- what the Python drop-in passes: staged upload inside the library
    for _ in range(1):
        ctx.run_pair(img1, img2, c1, r1, c2, r2, b, s, angles, 0.0)
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    pg_steps = max(2, min(e2e_steps, 5))
    for _ in range(pg_steps):
        ctx.run_pair(img1, img2, c1, r1, c2, r2, b, s, angles, 0.0)
    dt_pg = max_over_ranks(time.perf_counter() - t0)
    e2e_pageable = {"value": total_points * pg_steps / dt_pg, "unit": UNIT, "ms_per_step": 1e3 * dt_pg / pg_steps,
                    "steps": pg_steps, "call": "sid_run_pair with pageable NumPy images (staged through a pinned double buffer)"}

    # ---- north_star's split of ONE pair over the ranks (strong scaling): slab upload + NVLink all-gather + one result all-gather
    strong = None
    if world > 1 and not args.no_strong:
        from sea_ice_drift_b200 import sharding
        si1, si2, sc1, sr1, sc2, sr2, sb, _ = syn.make_config(args.workload, seed=0, side=args.side or None, grid=args.grid or None)
        si1 = torch.from_numpy(si1).pin_memory().numpy()
        si2 = torch.from_numpy(si2).pin_memory().numpy()
        table = None
        for _ in range(2):
            table = sharding.use_mcc_batch_split(sc1, sr1, sc2, sr2, sb, si1, si2, s, 0.0, angles=angles, device=local_rank)
        barrier(); torch.cuda.synchronize()
        st_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(st_steps):
            table = sharding.use_mcc_batch_split(sc1, sr1, sc2, sr2, sb, si1, si2, s, 0.0, angles=angles, device=local_rank)
        dt_st = max_over_ranks(time.perf_counter() - t0)
        plan = sharding.SplitPlan(si1.shape[0], si1.shape[1], world)
        strong = {"scaling": "strong", "value": len(sc1) * st_steps / dt_st, "unit": UNIT, "ms_per_pair": 1e3 * dt_st / st_steps,
                  "steps": st_steps, "points": int(len(sc1)),
                  "h2d_bytes_per_rank": int(2 * plan.rows_per * plan.cols), "nvlink_allgather_bytes_per_rank": int(2 * plan.gather_bytes),
                  "result_allgather_bytes": int(-(-len(sc1) // world) * world * 40),
                  "collectives_per_step": "2 in-place all_gather_into_tensor of the image row slabs (NCCL over NVLink) + 1 of the result rows",
                  "table_checksum": float(np.nansum(table[:, :2]))}
        if rank == 0 and world == int(os.environ.get("WORLD_SIZE", "1")):
            # the sharded table must equal this rank's own single-GPU table of the same pair
            ctx.set_pair(si1, si2)
            one = ctx.run(sc1, sr1, sc2, sr2, sb, s, angles, 0.0)
            strong["equals_single_gpu_table"] = bool(np.array_equal(table, one, equal_nan=True))
            ctx.set_pair(img1p, img2p)
        barrier()

    # ---- the other single-GPU BASELINE configurations at full size (device-resident)
    configs = None
    if world == 1 and not args.no_configs and args.workload == "cfg2":
        configs = {}
        ctx.set_stream(stream.cuda_stream)
        for name, steps_c in (("cfg3", 10), ("cfg4", 4), ("cfg1", 20)):
            cc = dict(syn.CONFIGS[name])
            if args.side:
                cc["side"] = args.side if name != "cfg1" else min(args.side, cc["side"])
            if args.grid:
                cc["grid"] = args.grid
            fg_note = None
            if name == "cfg1":
                i1, i2, q1, q2, q3, q4, q5, cc2 = syn.make_config("cfg1", seed=0, side=cc["side"], grid=cc["grid"])
                cpts = [q1, q2, q3, q4, q5]
                try:        # BASELINE configs[0] says "ORB first guess": borders from a real one (outside the timed region)
                    cpts = list(syn.orb_first_guess_inputs(i1, i2, cc["grid"], cc["img_size"]))
                    fg_note = ("ORB first guess (cv2.ORB + GPU Hamming matcher + prepare_first_guess): borders %d..%d, %.0f %% at %d"
                               % (cpts[4].min(), cpts[4].max(), 100.0 * (cpts[4] == cpts[4].min()).mean(), cpts[4].min()))
                except Exception as exc:        # no cv2 on the box: the seeded random borders of synthetic.make_config
                    fg_note = "random borders 20..50 (ORB first guess unavailable: %s)" % type(exc).__name__
                ctx.set_pair(i1, i2)
            else:       # cfg3 / cfg4 use the EW pair that is already resident (same seed, same warp)
                m = syn.rotation_matrix(img1.shape, cc["warp"][1])
                cpts = list(syn.hot_loop_inputs(img1, m, cc["grid"], cc["img_size"], cc["border"], rank))
            ms_c, k_ms, o_c, st_c = time_resident(ctx, torch, stream, dev, cpts, cc["img_size"], cc["angles"], steps_c, 3)
            fl = float(flops_per_point(cc["img_size"], cpts[4][st_c == 1], len(cc["angles"])).sum())
            configs[name] = {"workload": workload_description(name, cc, len(cpts[0])), "points": int(len(cpts[0])),
                             "valid": int((st_c == 1).sum()), "value": len(cpts[0]) / (ms_c * 1e-3), "unit": UNIT,
                             "ms_per_step": ms_c, "kernel_ms": k_ms, "steps": steps_c, "kernel": ctx.last_kernel_name,
                             "tflops_equiv": fl / (k_ms * 1e-3) / 1e12}
            if fg_note:
                configs[name]["first_guess"] = fg_note
        ctx.set_stream(None)
        # cfg5 (time series of EW pairs, 300 x 300 grid each) through sharding.use_mcc_series: every pair is uploaded inside
        # the timed region (pinned host images), one table per pair comes back; end-to-end wall clock.  One GPU's share of
        # the series: 4 pairs (the series cycles through the resident host pair; scratch/time_series.py runs all 16 at N > 1)
        try:
            from sea_ice_drift_b200 import sharding
            c5 = dict(syn.CONFIGS["cfg5"])
            if args.grid:
                c5["grid"] = args.grid
            m5 = syn.rotation_matrix(img1.shape, c5["warp"][1])
            p5 = list(syn.hot_loop_inputs(img1, m5, c5["grid"], c5["img_size"], c5["border"], rank))
            series = [(img1p, img2p) + tuple(p5)] * 4
            sharding.use_mcc_series(series[:2], c5["img_size"], 0.0, angles=c5["angles"])            # warm-up (allocations)
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                tabs = sharding.use_mcc_series(series, c5["img_size"], 0.0, angles=c5["angles"])
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            nvec = sum(len(t) for t in tabs)
            configs["cfg5"] = {"workload": "cfg5: time series, %d of 16 pairs %dx%d on this GPU, %d-point PM grid each (300x300 requested), img_size=%d, "
                                           "border=%s, angles=%s; sharding.use_mcc_series, every pair uploaded inside the timed region"
                                           % (len(series), img1.shape[0], img1.shape[1], len(p5[0]), c5["img_size"], c5["border"], c5["angles"]),
                               "points": int(nvec), "value": nvec / best, "unit": UNIT, "ms_per_pair": best * 1e3 / len(series),
                               "timing": "end to end, host clock, best of 3", "tflops_equiv": None}
        except Exception as exc:            # reported, never fatal for the headline line
            configs["cfg5"] = {"error": repr(exc)}
        ctx.set_pair(img1p, img2p)
    drop_in = None
    if ref_drop_in is not None:
        try:
            drop_in = drop_in_block(syn, ref_drop_in)
        except Exception as exc:            # reported, never fatal for the headline line
            drop_in = {"error": repr(exc)}
        ctx.set_pair(img1p, img2p)
    drop_in_ew = None
    if world == 1 and not args.no_configs and args.workload == "cfg2":
        try:
            drop_in_ew = drop_in_ew_block(syn, img1, img2, angles, grid=args.grid or 200)
        except Exception as exc:
            drop_in_ew = {"error": repr(exc)}
        ctx.set_pair(img1p, img2p)

    line = None
    if rank == 0:
        # ---- parity on a seeded sample + CPU baseline (rank 0, N = 1 only for the baseline)
        from oracle import c_oracle
        sel = np.sort(np.random.default_rng(123).choice(n, min(400, n), replace=False))
        exact, _ = c_oracle.use_mcc_batch(c1[sel], r1[sel], c2[sel], r2[sel], b[sel], img1, img2, s, 0.0, angles=angles)
        ok = ~np.isnan(exact[:, 0])
        parity = {"sample_points": int(len(sel)), "nan_pattern_equal": bool(np.array_equal(np.isnan(out[sel]), np.isnan(exact))),
                  "position_angle_equal": int((out[sel][ok, :3] == exact[ok, :3]).all(axis=1).sum()),
                  "r_bit_equal": int((out[sel][ok, 3] == exact[ok, 3]).sum()), "compared": int(ok.sum()),
                  "max_abs_dh": float(np.abs(out[sel][ok, 4] - exact[ok, 4]).max()) if ok.any() else 0.0,
                  "checker": "oracle/mcc_oracle.c (exact CPU restatement)", "e2e_equals_resident": same_as_resident}
        cpu = None
        if cpu_run is not None:
            threads, rate, m, dt, rows, csel = cpu_run
            agree = int((np.nan_to_num(rows[:, :3], nan=-1) == np.nan_to_num(out[csel][:, :3], nan=-1)).all(axis=1).sum())
            cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": cpu_kind(), "other_thread_counts": cpu_extra,
                   "source": "oracle/_ref: unmodified sea_ice_drift/pmlib.py use_mcc_mp through a fork Pool"
                             if cpu_kind() == "reference" else "oracle/pm_oracle.py (port)",
                   "sample": "%d seeded random grid points of the same workload in %.1f s (fork Pool, %d workers); "
                             "%d/%d rows agree with the GPU in position and angle" % (m, dt, threads, agree, m)}
            if cpu_kind() == "reference":
                parity["vs_reference"] = reference_parity(out[csel], rows, [x[csel] for x in (c1, r1, c2, r2, b)],
                                                          img1, img2, s, angles)
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        sm_max = float(peaks.get("sm_max_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0)
        peak_tflops = sm_count * 128 * 2 * sm_max * 1e6 / 1e12
        achieved = flops_step / (kernel_ms * 1e-3) / 1e12
        traffic = traffic_by_kernel = traffic_algorithmic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get(args.workload)
            traffic_by_kernel = tj.get(args.workload + "_by_kernel")
            traffic_algorithmic = tj.get(args.workload + "_algorithmic")
        except (OSError, ValueError):
            pass
        tensor_peak = float(peaks.get("bf16_tflops") or 1590.0)
        imma_peak = sm_count * 1950 * 2 * sm_max * 1e6 / 1e12
        i8_peak = sm_count * 7949 * 2 * sm_max * 1e6 / 1e12       # tcgen05 kind::i8 measured: profiles/r02_tcgen05_i8_rates.txt
        if configs:
            for cdict in configs.values():
                if cdict.get("tflops_equiv") is None:          # cfg5 is an end-to-end figure (uploads inside), no kernel roofline
                    continue
                cdict["roofline_frac"] = cdict["tflops_equiv"] / tensor_peak
                cdict["frac_of_fp32_fma_peak"] = cdict["tflops_equiv"] / peak_tflops
        # main object: the correlation of the dominant kernel runs on the tensor pipe (exact u8 x u8 -> s32 IMMA), so the
        # bounding roofline is "tensor" against the measured dense bf16 peak of MEASURED_PEAKS.json
        roofline = {"bound": "tensor", "achieved": achieved, "peak": tensor_peak, "unit": "TFLOP/s",
                    "frac": achieved / tensor_peak, "traffic": traffic,
                    "traffic_by_kernel": traffic_by_kernel, "traffic_algorithmic": traffic_algorithmic,
                    "traffic_note": "DRAM bytes of one step = both kernels (ncu --set full, profiles/traffic.json); the excess over the "
                                    "algorithmic bytes is the winning-map hand-off to pm_tail_kernel; HBM is < 4 % busy",
                    "kernel": kernel_name, "kernel_ms": kernel_ms, "step_ms": ms_per_step,
                    "kernels_per_step": "%s (correlation on tcgen05.mma kind::i8, dominant) + pm_tail_kernel (peak statistics)"
                                        % kernel_name.split("::")[-1],
                    "algorithmic_flops_per_launch": flops_step,
                    "peak_source": ("bf16_tflops of MEASURED_PEAKS.json (of measured)" if peaks.get("bf16_tflops")
                                    else "1.59 PFLOP/s (of fallback)"),
                    "u8_tcgen05_peak": i8_peak, "frac_of_u8_tcgen05_peak": achieved / i8_peak,
                    "u8_mma_sync_peak": imma_peak, "frac_of_u8_mma_sync_peak": achieved / imma_peak,
                    "frac_of_fp32_fma_peak": achieved / peak_tflops,
                    "note": "algorithmic FLOPs = sum over points and angles of 2 s^2 R^2 (SURVEY 8d), multiply-adds of the direct-form "
                            "correlation only; they run as tcgen05.mma kind::i8 (u8 x u8 -> s32, TMEM accumulators; measured pipe rate "
                            "7949 MAC/clk/SM = u8_tcgen05_peak, profiles/r02_tcgen05_i8_rates.txt). pm_ws_kernel (default for search "
                            "radii <= 22 at img_size 35) is a warp-specialised pipeline, one CTA of 28 warps per SM: the tensor pipe is ~15 % busy and no "
                            "pipe is saturated; the time is set by the instruction latency chains of the gather / window-statistics / "
                            "normalisation roles at 28 resident warps (ncu: 13 cycles per issued warp instruction, IPC 2.1; "
                            "profiles/r02_pm_ws_ncu_roles.txt). BASELINE.json's own figure, the fraction of the FP32-FMA roofline, is in "
                            "roofline_fma",
                    "hbm_gbs_measured_peak": peaks.get("hbm_gbs")}
        roofline_fma = {"bound": "fma", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                        "frac": achieved / peak_tflops, "traffic": traffic,
                        "peak_source": "FP32 FMA pipe: %d SMs x 128 lanes x 2 x %.0f MHz (sm_max_mhz of MEASURED_PEAKS.json)"
                                       % (sm_count, sm_max),
                        "note": "BASELINE.json's metric ('% of FP32 FMA roofline', SURVEY 8d); above 1 because the multiply-adds "
                                "run on the tensor pipe"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": workload_description(args.workload, cfg, n), "points_per_gpu": n,
                           "l2": "inputs larger than L2: 2 x %.0f MB image pair per GPU, no flush between steps"
                                 % (img1.nbytes / 1e6),
                           "parallelism": "grid points sharded per GPU, one pair per rank, no data-path collective",
                           "host_binding": ("rank bound to the %d cores local to its GPU (NVML affinity)" % len(bound)) if bound
                                           else "none"},
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "ms_per_step": 1e3 * dt_e2e / e2e_steps,
                        "call": "sid_run_pair: pinned host image pair + host point arrays in, host result table out; "
                                "upload in row bands overlapped with the fused kernel"},
                "e2e_pageable": e2e_pageable,
                "roofline": roofline, "roofline_fma": roofline_fma, "cpu_baseline": cpu, "parity": parity,
                "configs": configs, "strong": strong, "drop_in": drop_in, "drop_in_ew": drop_in_ew}
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    return 0


class JsonOnlyStdout(object):
    """Keep stdout clean for the ONE JSON line: while active, file descriptor 1 points at stderr (so banners
    printed by native libraries such as NCCL's version line go there); the JSON goes to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)
        return False


if __name__ == "__main__":
    a = parse_args()
    import contextlib
    import io
    buf = io.StringIO()
    with JsonOnlyStdout():
        with contextlib.redirect_stdout(buf):
            rc = (drop_in_reference_child(a.drop_in_reference) if a.drop_in_reference
                  else main_reference(a) if a.impl == "reference" else main_ours(a))
    line = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    if line:
        print(line[-1], flush=True)
    sys.exit(rc)
