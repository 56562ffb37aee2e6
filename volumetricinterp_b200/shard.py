"""Multi-GPU: records (fit) and query points (Estimate) shard with no exchange during compute
(the reference's record loop, interpolate.py:511, has no cross-record dependence).

One process per GPU (torchrun).  Each rank fits a contiguous block of records; the only collective is
ONE gather of the small per-record results (coefficients R x N, chi^2, lambda, rank, status) — over
NCCL/NVLink on GPUs, gloo in the CPU tests.  The covariance (R x N x N) is deliberately NOT gathered:
each rank keeps / writes its own block.
"""
import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous block [lo, hi) of n items for `rank` (sizes differ by at most one)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_rows(t, counts, group=None):
    """Concatenate per-rank tensors with differing first dimensions (counts[r] rows on rank r)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])


def fit_records_sharded(model, lat, lon, alt, value, error, reg_matrices=None, method='chi2', group=None,
                        fit_fn=None, **kw):
    """Every rank passes the SAME full (R, P) value/error (or at least its own rows); returns the
    gathered FitResult-like dict on every rank.  Covariance stays local (key 'Covariance_local')."""
    import torch
    import torch.distributed as dist
    from . import fit as _fit
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    R = value.shape[0]
    lo, hi = shard_bounds(R, world, rank)
    fn = fit_fn or _fit.fit_records
    res = fn(model, lat, lon, alt, value[lo:hi], error[lo:hi], reg_matrices, method, to_host=False, **kw)
    counts = [shard_bounds(R, world, r)[1] - shard_bounds(R, world, r)[0] for r in range(world)]
    out = {}
    for key in ("Coeffs", "chi_sq", "reg_params", "rank", "status"):
        t = getattr(res, key)
        t = t if isinstance(t, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(t))
        out[key] = all_gather_rows(t, counts, group)
    out["Covariance_local"] = res.Covariance
    out["local_rows"] = (lo, hi)
    return out
