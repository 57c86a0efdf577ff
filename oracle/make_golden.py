"""TEST INFRASTRUCTURE ONLY -- regenerate tests/golden/*.npz by running the
UNMODIFIED reference (``/root/reference/sea_ice_drift/pmlib.py`` through
oracle/ref_import.py) on small seeded inputs.  Only runs in the build container.

    python -m oracle.make_golden

The reference's own tests hold no numeric pins for this path (tests.py:296-346
assert shapes only), so these vectors are what pins the oracles and the GPU path.
"""
import contextlib
import io
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.ref_import import load_reference          # noqa: E402
from sea_ice_drift_b200 import synthetic as syn       # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

VARIANTS = {
    # name: (img_size, angles, alpha0, border spec, kwargs)
    "default": (35, [-3, 0, 3], 0.0, (20, 30), {}),
    "one_angle": (35, [0], 0.0, (20, 26), {}),
    "rot_order1": (35, [-3, 0, 3], 1.25, (20, 24), {"rot_order": 1}),
    "hes_smth": (35, [-3, 0, 3], 0.0, (20, 24), {"hes_smth": True}),
    "raw_hes_mcc_norm": (35, [-2, 0, 2], 0.0, (20, 24), {"hes_norm": False, "mcc_norm": True}),
    "even50_7angles": (50, [-3, -2, -1, 0, 1, 2, 3], -3.85, (20, 22), {}),
    "s51_b60": (51, [-3, 0, 3], 0.0, (60, 60), {}),
    "s21_b9": (21, [-4, 4], 0.0, (9, 12), {"rot_order": 1, "hes_smth": True}),
}


def main():
    pm = load_reference()
    warnings.simplefilter("ignore")
    os.makedirs(OUT, exist_ok=True)
    side = 420
    img1 = syn.speckle_image((side, side), seed=11)
    m = syn.rotation_matrix((side, side), 1.6)
    m[0, 2] += 4.0
    m[1, 2] -= 3.0
    img2 = syn.warp_pair(img1, m, seed=11)
    img1 = img1.copy()
    img1[150:171, 230:262] = 0                      # invalid patch -> NaN vectors (pmlib.py:152-154)
    data = {"img1": img1, "img2": img2, "matrix": m}
    names = []
    for vi, (name, (s, angles, alpha0, bspec, kw)) in enumerate(VARIANTS.items()):
        n_side = 5 if s == 51 else 9
        c1, r1, c2fg, r2fg, brd = syn.hot_loop_inputs(img1, m, n_side, s, bspec, seed=100 + vi,
                                                      fg_noise=1.5, inset=s + bspec[1] + 12)
        pm.shared_args = (c1, r1, c2fg, r2fg, brd, img1, img2, s, alpha0)
        pm.shared_kwargs = dict(angles=angles, **kw)
        with contextlib.redirect_stdout(io.StringIO()):
            rows = np.array([pm.use_mcc_mp(i) for i in range(len(c1))], dtype=np.float64)
        data.update({name + "/c1": c1, name + "/r1": r1, name + "/c2fg": c2fg, name + "/r2fg": r2fg,
                     name + "/border": brd, name + "/out": rows,
                     name + "/meta": np.array([s, alpha0] + list(angles), dtype=np.float64),
                     name + "/kw": np.array([kw.get("rot_order", 0), kw.get("hes_norm", True),
                                             kw.get("hes_smth", False), kw.get("mcc_norm", False)], dtype=np.int64)})
        names.append(name)
        print("%-18s points=%3d NaN rows=%d" % (name, len(c1), int(np.isnan(rows[:, 0]).sum())))
    data["variants"] = np.array(names)
    np.savez_compressed(os.path.join(OUT, "pm_points.npz"), **data)

    # per-stage vectors: get_template, matchTemplate, get_hessian, rotate_and_match
    rng = np.random.default_rng(5)
    stage = {"img": img1[:200, :220].copy()}
    tcases = []
    for k in range(16):
        s = int(rng.choice([35, 51, 50, 7]))
        order = k % 2
        c, r = [(rng.uniform(40, 180), rng.uniform(40, 160)), (float(rng.integers(40, 180)), float(rng.integers(40, 160))),
                (rng.integers(40, 180) + 0.5, rng.integers(40, 160) + 0.5), (rng.uniform(-3, 223), rng.uniform(-3, 203))][k % 4]
        ang = float(rng.choice([0, 3, -3, 30, 90, -10.5, 180]))
        tcases.append([c, r, ang, s, order])
        stage["tpl_%d" % k] = pm.get_template(stage["img"], c, r, ang, s, rot_order=order)
    stage["tpl_cases"] = np.array(tcases)
    import cv2
    for k in range(4):
        s = [35, 51, 50, 35][k]
        b = [20, 30, 20, 8][k]
        y, x = rng.integers(0, side - s - 2 * b - 1, 2)
        win = img2[y:y + s + 2 * b + (s % 2 == 0), x:x + s + 2 * b + (s % 2 == 0)]
        ty, tx = rng.integers(0, 200 - s, 2)
        tpl = img2[y + b - 1:y + b - 1 + s, x + b + 2:x + b + 2 + s] if k % 2 == 0 else img2[ty:ty + s, tx:tx + s]
        ccm = cv2.matchTemplate(win, np.ascontiguousarray(tpl), cv2.TM_CCOEFF_NORMED)
        stage["mt_win_%d" % k] = win.copy()
        stage["mt_tpl_%d" % k] = np.ascontiguousarray(tpl)
        stage["mt_out_%d" % k] = ccm
        for hn, hs in ((1, 0), (0, 0), (1, 1), (0, 1)):
            stage["hes_%d_%d%d" % (k, hn, hs)] = pm.get_hessian(ccm, hes_norm=bool(hn), hes_smth=bool(hs))
    # rotate_and_match with an explicit, non-square window and even template (cf. tests.py:336-337)
    ram = pm.rotate_and_match(img1, 210.3, 120.6, 50, img2[40:190, 130:300], -3.85, angles=[-3, -2, -1, 0, 1, 2, 3])
    stage["ram_scalars"] = np.array(ram[:5], dtype=np.float64)
    stage["ram_result"] = ram[5]
    stage["ram_template"] = ram[6]
    np.savez_compressed(os.path.join(OUT, "pm_stages.npz"), **stage)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
