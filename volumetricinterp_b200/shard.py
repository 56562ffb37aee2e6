"""Multi-GPU: records (fit) and query points (Estimate) shard with no exchange during compute
(the reference's record loop, interpolate.py:511, has no cross-record dependence, and
Estimate.__call__, estimate.py:75-123, none between query points).

One process per GPU (torchrun).  Each rank fits a contiguous block of records; the only collective is
ONE gather of the small per-record results (coefficients R x N, chi^2, lambda, rank, status) — over
NCCL/NVLink on GPUs, gloo in the CPU tests.  The covariance (R x N x N) is deliberately NOT gathered:
each rank keeps / writes its own block.  Estimate shards the query points the same way; every rank holds
the full (small) coefficient set and the outputs are gathered, or left per shard for the caller to write.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous block [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(t, counts, group=None):
    """Concatenate per-rank tensors with differing first dimensions (counts[r] rows on rank r; zero allowed)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mx = max(max(counts), 1)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def _row_counts(local, group):
    """Rows held by every rank (one tiny all-gather)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend(group) == 'nccl' else torch.device('cpu')
    mine = torch.tensor([int(local)], dtype=torch.int64, device=dev)
    bufs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(bufs, mine, group=group)
    return [int(b.item()) for b in bufs]


def fit_records_sharded(model, lat, lon, alt, value, error, reg_matrices=None, method='chi2', group=None,
                        fit_fn=None, presharded=False, **kw):
    """Record-sharded fit.  presharded=False: every rank passes the SAME full (R, P) value/error and fits its
    contiguous block; presharded=True: value/error already hold only this rank's records (ranks read or
    generate their own slice of the file).  Returns the gathered per-record results on every rank (a dict of
    tensors, records in rank order); the covariance stays local: 'Covariance_local' (host pinned view with
    to_host=True, device tensor otherwise), rows 'local_rows' of the gathered arrays.  A rank may hold zero
    records (R < world size, or an empty time window)."""
    import torch
    import torch.distributed as dist
    from . import fit as _fit
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    fn = fit_fn or _fit.fit_records
    if presharded:
        counts = _row_counts(value.shape[0], group)
        lo = sum(counts[:rank])
        hi = lo + counts[rank]
        v, e = value, error
    else:
        R = value.shape[0]
        lo, hi = shard_bounds(R, world, rank)
        counts = [shard_bounds(R, world, r)[1] - shard_bounds(R, world, r)[0] for r in range(world)]
        v, e = value[lo:hi], error[lo:hi]
    kw.setdefault('to_host', False)
    res = fn(model, lat, lon, alt, v, e, reg_matrices, method, **kw)
    small = getattr(res, 'device_small', None) or {}
    out = {}
    for key in ("Coeffs", "chi_sq", "reg_params", "rank", "status"):
        t = small.get(key, getattr(res, key))
        t = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t))
        out[key] = all_gather_rows(t, counts, group)
    out["Covariance_local"] = res.Covariance
    out["local_rows"] = (lo, hi)
    out["local"] = res
    return out


def estimate_sharded(est, C, lat, lon, alt, check_hull=True, group=None, gather=True):
    """Point-sharded Estimate: C (Rsel, N) and lat/lon/alt (npts,) are the same CUDA tensors on every rank;
    rank r evaluates its contiguous block of points for all Rsel records (`Estimate.evaluate_device`).  Returns
    (out, (lo, hi)): with gather=True `out` is the full (Rsel, npts) tensor on every rank (one all-gather of the
    per-shard outputs), otherwise the local (Rsel, hi - lo) block."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    npts = lat.numel()
    lo, hi = shard_bounds(npts, world, rank)
    local = est.evaluate_device(C, lat[lo:hi].contiguous(), lon[lo:hi].contiguous(), alt[lo:hi].contiguous(), check_hull)
    if not gather:
        return local, (lo, hi)
    counts = [shard_bounds(npts, world, r)[1] - shard_bounds(npts, world, r)[0] for r in range(world)]
    full = all_gather_rows(local.t().contiguous(), counts, group)        # points are the ragged dimension
    return full.t().contiguous(), (lo, hi)
