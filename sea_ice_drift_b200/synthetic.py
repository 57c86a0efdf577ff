"""Seeded synthetic SAR-like image pairs with a known drift field, plus a minimal
Nansat stand-in, for tests and ``bench.py`` (there is no network for real
Sentinel-1 scenes).  Workload shapes follow BASELINE.json ``configs`` / SURVEY.md
section 8(d).  Pixel value 0 is reserved for "invalid", as the reference's
``get_uint8_image`` does (reference sea_ice_drift/lib.py:52-57)."""
import numpy as np
import cv2


def speckle_image(shape, seed, sigma=1.7):
    """uint8 texture in [1, 255]: gamma(4) speckle, blurred, 1..99 % stretched."""
    rng = np.random.default_rng(seed)
    field = rng.gamma(4.0, 1.0, size=shape).astype(np.float32)
    field = cv2.GaussianBlur(field, (0, 0), sigma)
    sample = field[::7, ::7]
    lo, hi = np.percentile(sample, [1.0, 99.0])
    scaled = 1.0 + 254.0 * (field - lo) / (hi - lo)
    return np.clip(scaled, 1.0, 255.0).astype(np.uint8)


def warp_pair(img1, matrix, seed, noise_sd=6.0):
    """img2 = img1 moved by the 2x3 affine ``matrix`` (img1 px -> img2 px) plus
    independent noise; pixels with no source stay 0 (invalid)."""
    h, w = img1.shape
    moved = cv2.warpAffine(img1, np.asarray(matrix, np.float64), (w, h), flags=cv2.INTER_LINEAR,
                           borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    rng = np.random.default_rng(seed + 7919)
    out = np.empty_like(moved)
    step = 1024
    for y in range(0, h, step):
        blk = moved[y:y + step].astype(np.float32)
        noisy = blk + rng.normal(0.0, noise_sd, size=blk.shape).astype(np.float32)
        noisy = np.clip(np.rint(noisy), 1.0, 255.0)
        noisy[blk == 0] = 0
        out[y:y + step] = noisy.astype(np.uint8)
    return out


def shift_matrix(dx, dy):
    return np.array([[1.0, 0.0, dx], [0.0, 1.0, dy]])


def rotation_matrix(shape, degrees):
    cy, cx = (shape[0] - 1) / 2.0, (shape[1] - 1) / 2.0
    a = np.radians(degrees)
    ca, sa = np.cos(a), np.sin(a)
    return np.array([[ca, -sa, cx - ca * cx + sa * cy], [sa, ca, cy - sa * cx - ca * cy]])


def apply_affine(matrix, x, y):
    m = np.asarray(matrix)
    return m[0, 0] * x + m[0, 1] * y + m[0, 2], m[1, 0] * x + m[1, 1] * y + m[1, 2]


def grid_points(shape, n_side, inset):
    xs = np.linspace(inset, shape[1] - 1 - inset, n_side)
    ys = np.linspace(inset, shape[0] - 1 - inset, n_side)
    gx, gy = np.meshgrid(xs, ys)
    return gx.ravel(), gy.ravel()


def hot_loop_inputs(img1, matrix, n_side, img_size, border, seed, fg_noise=2.0, inset=None,
                    subpixel=True):
    """Arrays the per-point loop consumes (the argument tuple of the reference's
    ``_init_pool``, pmlib.py:438): c1, r1 (float px on img1), c2fg, r2fg
    (integer-valued float px on img2), border (integer-valued float).
    ``border`` may be an int or a (lo, hi) range drawn per point."""
    rng = np.random.default_rng(seed + 104729)
    bmax = border if np.isscalar(border) else border[1]
    if inset is None:
        inset = int(img_size + bmax + 3 * fg_noise + 16)
    c1, r1 = grid_points(img1.shape, n_side, inset)
    if subpixel:
        c1 = c1 + rng.uniform(-0.5, 0.5, c1.size)
        r1 = r1 + rng.uniform(-0.5, 0.5, r1.size)
    else:
        c1, r1 = np.round(c1), np.round(r1)
    tx, ty = apply_affine(matrix, c1, r1)
    c2fg = np.round(tx + rng.normal(0, fg_noise, tx.size))
    r2fg = np.round(ty + rng.normal(0, fg_noise, ty.size))
    if np.isscalar(border):
        brd = np.full(c1.size, float(border))
    else:
        brd = np.floor(rng.uniform(border[0], border[1] + 1, c1.size))
    hws = img_size // 2 + 1
    ok = ((c2fg - brd - hws > 0) & (r2fg - brd - hws > 0) &
          (c2fg + brd + hws < img1.shape[1]) & (r2fg + brd + hws < img1.shape[0]))
    return c1[ok], r1[ok], c2fg[ok], r2fg[ok], brd[ok]


CONFIGS = {
    # name: (image side, warp, grid side, img_size, border, angles)
    "cfg1": dict(side=2000, warp=("shift", 12.0, 12.0), grid=50, img_size=35, border=(20, 50), angles=[0]),
    "cfg2": dict(side=10400, warp=("rot", 2.0), grid=200, img_size=35, border=20, angles=[-3, 0, 3]),
    "cfg3": dict(side=10400, warp=("rot", 2.0), grid=200, img_size=35, border=20, angles=list(range(-10, 11))),
    "cfg4": dict(side=10400, warp=("rot", 2.0), grid=400, img_size=51, border=100, angles=[-3, 0, 3]),
    "cfg5": dict(side=10400, warp=("rot", 2.0), grid=300, img_size=35, border=20, angles=[-3, 0, 3], pairs=16),
}


def make_config(name, seed=0, side=None, grid=None):
    """Build (img1, img2, c1, r1, c2fg, r2fg, brd, meta) for a BASELINE config;
    ``side``/``grid`` shrink it for tests."""
    cfg = dict(CONFIGS[name])
    if side:
        cfg["side"] = side
    if grid:
        cfg["grid"] = grid
    shape = (cfg["side"], cfg["side"])
    img1 = speckle_image(shape, seed)
    if cfg["warp"][0] == "shift":
        m = shift_matrix(cfg["warp"][1], cfg["warp"][2])
    else:
        m = rotation_matrix(shape, cfg["warp"][1])
    img2 = warp_pair(img1, m, seed)
    pts = hot_loop_inputs(img1, m, cfg["grid"], cfg["img_size"], cfg["border"], seed)
    cfg["matrix"] = m
    return (img1, img2) + pts + (cfg,)


def orb_matches(img1, img2, n_features=20000, ratio_test=0.6):
    """Feature-tracking matches x1, y1, x2, y2 (pixels) for BASELINE configs[0] ("ORB first guess"): cv2.ORB keypoints with
    the reference's detector settings (ftlib.py:26-62), this package's GPU Hamming matcher + Lowe ratio test
    (``ftlib.get_match_coords``), and an outlier rejection (the reference's lstsq_filter, ftlib.py:203-234, fits a plane
    to the displacements; the synthetic drift is smooth, so a median test does the same job here)."""
    import cv2
    from . import ftlib
    cv2.setRNGSeed(0)
    det = cv2.ORB_create()
    det.setEdgeThreshold(34); det.setMaxFeatures(n_features); det.setNLevels(7); det.setPatchSize(34)
    kp1, d1 = det.detectAndCompute(img1, None)
    kp2, d2 = det.detectAndCompute(img2, None)
    x1, y1, x2, y2 = ftlib.get_match_coords(kp1, d1, kp2, d2, ratio_test=ratio_test)
    du, dv = x2 - x1, y2 - y1
    keep = (np.abs(du - np.median(du)) < 30.0) & (np.abs(dv - np.median(dv)) < 30.0)
    return x1[keep], y1[keep], x2[keep], y2[keep]


def orb_first_guess_inputs(img1, img2, n_side, img_size, inset=100, matches=None):
    """BASELINE configs[0] as the reference runs it: the per-point arrays of the hot loop from a real ORB first guess --
    ``orb_matches`` then ``pmlib.prepare_first_guess`` with the reference's default borders (20 ... 50 px by the distance to
    the nearest keypoint, pmlib.py:249-324).  Returns c1, r1, c2fg, r2fg, border of the grid points whose windows lie
    inside the images (the reference's ``gpi`` mask, pmlib.py:419-426)."""
    from . import pmlib
    x1, y1, x2, y2 = matches if matches is not None else orb_matches(img1, img2)
    n1, n2 = ArrayDomain(img1), ArrayDomain(img2)
    gx, gy = np.meshgrid(np.linspace(inset, img1.shape[1] - inset, n_side), np.linspace(inset, img1.shape[0] - inset, n_side))
    c1, r1 = gx.ravel(), gy.ravel()
    c2fg, r2fg, brd = pmlib.prepare_first_guess(c1, r1, n1, x1, y1, n2, x2, y2, img_size)
    hws = round(img_size / 2) + 1
    hyp = np.hypot(hws, hws)
    with np.errstate(invalid="ignore"):
        ok = (np.isfinite(c2fg) & np.isfinite(r2fg) & (c2fg - brd - hws > 0) & (r2fg - brd - hws > 0) &
              (c2fg + brd + hws < img2.shape[1]) & (r2fg + brd + hws < img2.shape[0]) &
              (c1 - hyp > 0) & (r1 - hyp > 0) & (c1 + hyp < img1.shape[1]) & (r1 + hyp < img1.shape[0]))
    return c1[ok], r1[ok], c2fg[ok], r2fg[ok], brd[ok]


class ArrayDomain(object):
    """Smallest object ``pattern_matching`` accepts in place of a Nansat: a uint8
    raster with an affine pixel<->lon/lat geolocation (duck type per SURVEY 8b:
    ``n[1]``, ``n.shape()``, ``n.transform_points``, ``n.get_corners``)."""

    def __init__(self, array, lon0=0.0, lat0=80.0, dlon=1e-3, dlat=-1e-3, rot_deg=0.0):
        self.array = np.ascontiguousarray(array, dtype=np.uint8)
        a = np.radians(rot_deg)
        self._fwd = np.array([[dlon * np.cos(a), -dlat * np.sin(a), lon0],
                              [dlon * np.sin(a), dlat * np.cos(a), lat0]])
        lin = self._fwd[:, :2]
        self._inv = np.linalg.inv(lin)

    def __getitem__(self, band):
        return self.array

    def shape(self):
        return self.array.shape

    def transform_points(self, x, y, DstToSrc=0, dst_srs=None):
        x = np.asarray(x, dtype=np.float64)
        y = np.asarray(y, dtype=np.float64)
        if DstToSrc == 0:
            lon = self._fwd[0, 0] * x + self._fwd[0, 1] * y + self._fwd[0, 2]
            lat = self._fwd[1, 0] * x + self._fwd[1, 1] * y + self._fwd[1, 2]
            return lon, lat
        dx, dy = x - self._fwd[0, 2], y - self._fwd[1, 2]
        return self._inv[0, 0] * dx + self._inv[0, 1] * dy, self._inv[1, 0] * dx + self._inv[1, 1] * dy

    def affine_maps(self, dst_srs=None):
        """(pixel -> destination x/y, pixel -> lon/lat) as 2 x 3 row-major matrices ``m0*col + m1*row + m2``: lets
        ``pattern_matching`` run its post-processing on the device.  This stand-in has no projection machinery: the
        destination SRS is lon/lat whatever ``dst_srs`` says (like ``transform_points``)."""
        return self._fwd.copy(), self._fwd.copy()

    def get_corners(self):
        h, w = self.array.shape
        cols = np.array([0, 0, w, w], dtype=np.float64)
        rows = np.array([0, h, 0, h], dtype=np.float64)
        return self.transform_points(cols, rows)
