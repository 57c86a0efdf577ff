"""Split tail for large maps with bigger CTAs: mixed borders at EW size, device time per step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sea_ice_drift_b200 import synthetic as syn, _lib
img1, img2, c1, r1, c2, r2, b, cfg = syn.make_config("cfg2", seed=0)
rng = np.random.default_rng(7); side = img1.shape[0]; nk = 50000
m = syn.rotation_matrix(img1.shape, 2.0)
kx, ky = rng.uniform(40, side - 40, nk), rng.uniform(40, side - 40, nk)
k2x, k2y = syn.apply_affine(m, kx, ky); k2x, k2y = k2x + rng.normal(0, 0.8, nk), k2y + rng.normal(0, 0.8, nk)
pts = list(syn.orb_first_guess_inputs(img1, img2, 200, 35, inset=150, matches=(kx, ky, k2x, k2y)))
ctx = _lib.Context(0); ctx.set_pair(img1, img2)
dev = torch.device("cuda", 0); stream = torch.cuda.current_stream(); ctx.set_stream(stream.cuda_stream)
def timed(p, label, env=None, angles=(-3, 0, 3)):
    if env: os.environ.update(env)
    d = torch.from_numpy(np.stack(p)).to(dev); o = torch.empty((len(p[0]), 5), dtype=torch.float64, device=dev)
    step = lambda: ctx.run_device(len(p[0]), *[d[k].data_ptr() for k in range(5)], int(p[4].max()), 35, list(angles), 0.0, o.data_ptr())
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): step()
    e1.record(stream); torch.cuda.synchronize()
    print("%-72s %7.3f ms / step   (%s)" % (label, e0.elapsed_time(e1) / 10, ctx.last_kernel_name), flush=True)
    if env:
        for k in env: del os.environ[k]
    return o.cpu().numpy()
a = timed(pts, "mixed borders, default (large maps: fused tail)")
for kb, th in ((100, 512), (100, 256), (100, 128), (64, 256)):
    r = timed(pts, "mixed borders, split tail up to %d KB, %d tail threads" % (kb, th), {"SID_PM_TAIL_SMEM_KB": str(kb), "SID_PM_TAIL_THREADS": str(th)})
    print("   same table:", np.array_equal(a, r, equal_nan=True))
cfg2pts = [c1, r1, c2, r2, b]
timed(cfg2pts, "cfg2 (uniform border 20), default tail (128 threads)")
timed(cfg2pts, "cfg2, 256 tail threads", {"SID_PM_TAIL_THREADS": "256"})
b30 = [c1, r1, c2, r2, np.full(len(c1), 30.0)]
x = timed(b30, "uniform border 30 (R = 61: split tail, default CTA size)")
y = timed(b30, "uniform border 30, 128 tail threads", {"SID_PM_TAIL_THREADS": "128"})
z = timed(b30, "uniform border 30, 512 tail threads", {"SID_PM_TAIL_THREADS": "512"})
print("   same table:", np.array_equal(x, y, equal_nan=True), np.array_equal(x, z, equal_nan=True))
