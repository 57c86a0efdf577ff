"""Grid points across the GPUs of one box: every point is independent
(reference pmlib.py:229-242 reads only shared read-only inputs), so ranks take
disjoint, equally heavy shares and the (N, 5) result table is assembled with one
all-gather.  No collective runs during the compute.

One process per GPU, launched by torchrun; ``torch.distributed`` only does the
plumbing (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def _dist():
    try:
        import torch.distributed as dist
    except ImportError:
        return None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist
    return None


def shard_indices(border, world_size, rank):
    """Indices of the points rank ``rank`` computes: points sorted by search radius
    (work per point ~ (2*border+1)^2) and dealt round-robin, so every rank gets the
    same mix of cheap and expensive points."""
    border = np.asarray(border, dtype=np.float64)
    order = np.argsort(-border, kind='stable')
    return np.sort(order[rank::world_size])


def gather_rows(local_rows, local_idx, n_total, dist, device=None):
    """All-gather the per-rank (n_r, 5) tables (padded to equal length) and scatter
    them back to the original point order.  One collective."""
    import torch
    world = dist.get_world_size()
    n_max = -(-n_total // world)
    pack = torch.full((n_max, 6), float('nan'), dtype=torch.float64)
    pack[:len(local_idx), 0] = torch.from_numpy(np.asarray(local_idx, dtype=np.float64))
    pack[:len(local_idx), 1:] = torch.from_numpy(np.ascontiguousarray(local_rows, dtype=np.float64))
    if device is not None:
        pack = pack.to(device)
    gathered = torch.empty((world * n_max, 6), dtype=torch.float64, device=pack.device)
    dist.all_gather_into_tensor(gathered, pack)
    g = gathered.cpu().numpy()
    keep = ~np.isnan(g[:, 0])
    out = np.full((n_total, 5), np.nan, dtype=np.float64)
    out[g[keep, 0].astype(np.int64)] = g[keep, 1:]
    return out


def use_mcc_batch_sharded(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, compute=None, **kwargs):
    """``use_mcc_batch`` over all ranks of an initialised process group (falls back to
    a plain single-GPU call when there is none).  Every rank passes the same inputs
    and receives the full table.  ``compute`` is the per-shard function (the GPU
    batch by default; the CPU tests inject a stand-in)."""
    if compute is None:
        from .pmlib import use_mcc_batch as compute
    dist = _dist()
    n = len(c1)
    if dist is None or n == 0:
        return compute(c1, r1, c2fg, r2fg, border, img1, img2, img_size, alpha0, **kwargs) if n else np.zeros((0, 5))
    rank, world = dist.get_rank(), dist.get_world_size()
    idx = shard_indices(border, world, rank)
    arrs = [np.asarray(a, dtype=np.float64)[idx] for a in (c1, r1, c2fg, r2fg, border)]
    local = compute(*arrs, img1, img2, img_size, alpha0, **kwargs) if len(idx) else np.zeros((0, 5))
    device = None
    if dist.get_backend() == 'nccl':
        import torch
        device = torch.device('cuda', torch.cuda.current_device())
    return gather_rows(local, idx, n, dist, device)
