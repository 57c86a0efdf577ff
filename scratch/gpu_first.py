import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sea_ice_drift_b200 import synthetic as syn, _lib
from oracle import c_oracle as co

def compare(name, got, ref):
    nan_eq = np.array_equal(np.isnan(got), np.isnan(ref))
    ok = ~np.isnan(ref[:, 0]) & ~np.isnan(got[:, 0])
    d = got[ok] - ref[ok]
    pos = int((np.abs(d[:, :2]).max(1) > 0).sum()) if ok.any() else 0
    ang = int((d[:, 2] != 0).sum()) if ok.any() else 0
    r_exact = int((d[:, 3] == 0).sum()); 
    print("%-34s n=%5d nan_eq=%s n_nan=%d pos_mismatch=%d ang_mismatch=%d r_exact=%d/%d max|dr|=%.3g max|dh|=%.3g maxrel|dh|=%.3g" % (
        name, len(ref), nan_eq, int(np.isnan(ref[:, 0]).sum()), pos, ang, r_exact, int(ok.sum()),
        np.abs(d[:, 3]).max() if ok.any() else 0, np.abs(d[:, 4]).max() if ok.any() else 0,
        (np.abs(d[:, 4]) / (1 + np.abs(ref[ok, 4]))).max() if ok.any() else 0), flush=True)
    return nan_eq and pos == 0 and ang == 0

ctx = _lib.Context(0)
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config('cfg2', seed=0, side=1500, grid=30)
img1z = img1.copy(); img1z[700:760, 700:760] = 0      # invalid patch -> NaN points
ctx.set_pair(img1z, img2)
allok = True
cases = [
    ("s35 b20 3ang", dict(s=35, angles=[-3, 0, 3], border=b)),
    ("s35 b20..50 3ang", dict(s=35, angles=[-3, 0, 3], border=np.floor(np.random.default_rng(1).uniform(20, 51, len(b))))),
    ("s35 b20 1ang", dict(s=35, angles=[0], border=b)),
    ("s35 b20 21ang", dict(s=35, angles=list(range(-10, 11)), border=b, n=200)),
    ("s35 b20 rot_order1", dict(s=35, angles=[-3, 0, 3], border=b, rot_order=1)),
    ("s35 smth", dict(s=35, angles=[-3, 0, 3], border=b, hes_smth=True)),
    ("s35 nonorm mccnorm", dict(s=35, angles=[-3, 0, 3], border=b, hes_norm=False, mcc_norm=True)),
    ("s50 (even) b20", dict(s=50, angles=[-3, 0, 3], border=b)),
    ("s34 (even) b23 mccnorm", dict(s=34, angles=[-2, 2], border=b + 3, mcc_norm=True)),
    ("s51 b100", dict(s=51, angles=[-3, 0, 3], border=b * 5, n=150)),
    ("s21 b10 (generic NW)", dict(s=21, angles=[-3, 0, 3], border=b / 2)),
    ("s64 b30 (generic NW)", dict(s=64, angles=[0, 5], border=b + 10, n=300)),
]
for name, kw in cases:
    s = kw['s']; n = kw.get('n', len(c1)); brd = np.asarray(kw['border'])[:n]
    opts = dict(rot_order=kw.get('rot_order', 0), hes_norm=kw.get('hes_norm', True), hes_smth=kw.get('hes_smth', False), mcc_norm=kw.get('mcc_norm', False))
    flags = _lib.flags_from_kwargs(opts['hes_norm'], opts['hes_smth'], opts['mcc_norm'])
    t0 = time.time(); got, st = ctx.run(c1[:n], r1[:n], c2[:n], r2[:n], brd, s, kw['angles'], 1.5, opts['rot_order'], flags, want_status=True); tg = time.time() - t0
    t0 = time.time(); ref, st2 = co.use_mcc_batch(c1[:n], r1[:n], c2[:n], r2[:n], brd, img1z, img2, s, 1.5, angles=kw['angles'], **opts); tc = time.time() - t0
    ok = compare(name, got, ref) and np.array_equal(st, st2)
    print("   gpu %.1f ms, C oracle %.0f ms, status_eq=%s" % (tg * 1e3, tc * 1e3, np.array_equal(st, st2)), flush=True)
    allok &= ok
# timing on the resident pair
import ctypes
for reps in range(2):
    t0 = time.time(); got = ctx.run(c1, r1, c2, r2, b, 35, [-3, 0, 3], 0.0); dt = time.time() - t0
    print("sid_run 900 pts: %.2f ms -> %.0f vec/s" % (dt * 1e3, len(c1) / dt))
print("ALL OK" if allok else "MISMATCH")
