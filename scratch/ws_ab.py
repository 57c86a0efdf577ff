"""A/B of the warp-specialised kernel (SID_PM_PATH=ws) against the mma.sync kernel (bit-identical outputs expected) + timing.
usage: python scratch/ws_ab.py [tiny|small|full]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sea_ice_drift_b200 import _lib, synthetic as syn


def run(ctx, path, args, reps=1):
    os.environ["SID_PM_PATH"] = path
    out = None
    ms = []
    for _ in range(reps):
        out, st = ctx.run(*args, want_status=True)
        ms.append(ctx.last_kernel_ms)
    return out, st, min(ms)


def case(ctx, name, img1, img2, pts, s, angles, **kw):
    ctx.set_pair(img1, img2)
    args = list(pts) + [s, angles, 1.5]
    flags = _lib.flags_from_kwargs(kw.get("hes_norm", True), kw.get("hes_smth", False), kw.get("mcc_norm", False))
    args += [kw.get("rot_order", 0), flags]
    ref, st_ref, ms_ref = run(ctx, "imma", args, kw.get("reps", 1))
    got, st, ms = run(ctx, "ws", args, kw.get("reps", 1))
    same = np.array_equal(got, ref, equal_nan=True) and np.array_equal(st, st_ref)
    neq = ~((got == ref) | (np.isnan(got) & np.isnan(ref)))
    nbad = int(neq.any(axis=1).sum())
    print("%-28s n=%6d valid=%6d  ws %.3f ms  imma %.3f ms  identical=%s (%d rows differ; per column %s)" %
          (name, len(pts[0]), int((st == 1).sum()), ms, ms_ref, same, nbad, neq.sum(axis=0).tolist()), flush=True)
    if not same:
        bad = np.nonzero(neq.any(axis=1))[0][:6]
        for i in bad:
            print("   row", i, "ws", got[i], "imma", ref[i], "border", pts[4][i], "st", st[i], st_ref[i])
    return same


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "small"
    ctx = _lib.Context(0)
    ok = True
    img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=2, side=1500, grid=26)
    img1 = img1.copy(); img1[690:770, 690:770] = 0
    rng = np.random.default_rng(9)
    n = len(c1)
    def pts(lo, hi, m=None):
        m = m or n
        brd = np.floor(rng.uniform(lo, hi + 1, n))[:m]
        return [x[:m].copy() for x in (c1, r1, c2, r2)] + [brd]
    ok &= case(ctx, "s35 a3 b20 (8 pts)", img1, img2, pts(20, 20, 8), 35, [-3, 0, 3])
    if mode == "tiny":
        return 0 if ok else 1
    ok &= case(ctx, "s35 a3 b20", img1, img2, pts(20, 20), 35, [-3, 0, 3])
    ok &= case(ctx, "s35 a1 b20", img1, img2, pts(20, 20), 35, [0])
    ok &= case(ctx, "s35 a2 b20-22", img1, img2, pts(20, 22), 35, [-3, 3], hes_norm=False, mcc_norm=True)
    ok &= case(ctx, "s35 a21 b20-24", img1, img2, pts(20, 24, 150), 35, list(range(-10, 11)))
    ok &= case(ctx, "s35 a7 b10-24", img1, img2, pts(10, 24), 35, [-9, -6, -3, 0, 3, 6, 9])
    ok &= case(ctx, "s50 a3 b10-20  (even)", img1, img2, pts(10, 20), 50, [-3, 0, 3])
    ok &= case(ctx, "s34 a2 b23", img1, img2, pts(23, 23), 34, [-2, 2], mcc_norm=True)
    ok &= case(ctx, "s21 a3 b8-14", img1, img2, pts(8, 14), 21, [-3, 0, 3])
    ok &= case(ctx, "s64 a2 b12-20", img1, img2, pts(12, 20, 200), 64, [0, 5])
    ok &= case(ctx, "s9 a1 b3-6", img1, img2, pts(3, 6), 9, [0])
    ok &= case(ctx, "s35 a3 b20-24 ord1 smth", img1, img2, pts(20, 24), 35, [-3, 0, 3], rot_order=1, hes_smth=True)
    ok &= case(ctx, "s35 a3 same angle x3", img1, img2, pts(20, 20), 35, [2, 2, 2])
    if mode == "full":
        for name in ("cfg2", "cfg3"):
            t0 = time.time()
            img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config(name, seed=0)
            print("built %s in %.1f s" % (name, time.time() - t0), flush=True)
            ok &= case(ctx, name + " full", img1, img2, [c1, r1, c2, r2, b], cfg["img_size"], cfg["angles"], reps=3)
    print("ALL IDENTICAL" if ok else "MISMATCHES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
