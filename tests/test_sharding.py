"""N > 1 path on the CPU: world_size-2 gloo run of the record sharding + the single gather
(SURVEY.md §8-e).  The per-rank fit is replaced by a deterministic stand-in (the product has no CPU
fit path); what is tested is the partition, the ragged all-gather and the ordering of the result."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_bounds_cover_everything():
    from volumetricinterp_b200.shard import shard_bounds
    for n in (0, 1, 7, 10000, 100003):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, R, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    from volumetricinterp_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = 5
    value = np.arange(R * 3, dtype=np.float64).reshape(R, 3)

    def fake_fit(model, lat, lon, alt, v, e, regs, method, to_host=False, **kw):
        r = v.shape[0]
        C = torch.from_numpy(v[:, :1] * np.arange(1, N + 1)[None, :])       # row-identifying coefficients
        return SimpleNamespace(Coeffs=C, Covariance=None, chi_sq=torch.from_numpy(v[:, 0].copy()),
                               reg_params=torch.from_numpy(v[:, 1:2].copy()),
                               rank=torch.full((r,), rank, dtype=torch.int32),
                               status=torch.zeros((r,), dtype=torch.int32))

    out = shard.fit_records_sharded(None, None, None, None, value, value, None, fit_fn=fake_fit)
    ok = (np.array_equal(out["Coeffs"].numpy(), value[:, :1] * np.arange(1, N + 1)[None, :])
          and np.array_equal(out["chi_sq"].numpy(), value[:, 0])
          and out["Coeffs"].shape == (R, N))
    lo, hi = out["local_rows"]
    owner = out["rank"].numpy()
    ok = ok and (owner[lo:hi] == rank).all()
    q.put((rank, bool(ok), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("R", [7, 10])
def test_two_rank_gather_gloo(R):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + R) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, R, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    bounds = sorted(b for _, _, b in res)
    assert bounds[0][0] == 0 and bounds[0][1] == bounds[1][0] and bounds[1][1] == R


def _worker_empty(rank, world, port, q):
    """R < world: one rank holds zero records (the gather must still complete, nothing may hang)."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from types import SimpleNamespace
    from volumetricinterp_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N = 4
    value = np.arange(3, dtype=np.float64).reshape(1, 3) + 5.0          # ONE record, two ranks

    def fake_fit(model, lat, lon, alt, v, e, regs, method, to_host=False, **kw):
        r = v.shape[0]
        return SimpleNamespace(Coeffs=torch.from_numpy(np.repeat(v[:, :1], N, axis=1).reshape(r, N)), Covariance=None,
                               chi_sq=torch.from_numpy(v[:, 0].copy()), reg_params=torch.from_numpy(v[:, 1:2].copy()),
                               rank=torch.full((r,), rank, dtype=torch.int32), status=torch.zeros((r,), dtype=torch.int32))

    out = shard.fit_records_sharded(None, None, None, None, value, value, None, fit_fn=fake_fit)
    ok = out["Coeffs"].shape == (1, N) and float(out["chi_sq"][0]) == 5.0 and out["local_rows"] in ((0, 1), (1, 1))
    # presharded: rank 1 brings nothing
    mine = value if rank == 0 else value[:0]
    out2 = shard.fit_records_sharded(None, None, None, None, mine, mine, None, fit_fn=fake_fit, presharded=True)
    ok = ok and out2["Coeffs"].shape == (1, N) and out2["local_rows"] == ((0, 1) if rank == 0 else (1, 1))
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_more_ranks_than_records_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_empty, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_pinned_result_buffers_are_leased_not_copied():
    """fit._pinned_buffer: the numpy views handed to the caller keep their pinned buffer checked out; it returns to
    the free list only when the last view dies, and is then re-used (no second host copy, no re-pinning)."""
    import gc
    import torch
    from volumetricinterp_b200 import fit
    real = fit._alloc_pinned

    def unpinned(n, dt):                       # no CUDA in the CPU test container: same pool logic, plain memory
        fit._NP_OF.update({torch.float64: np.float64, torch.int32: np.int32})
        return torch.empty((n,), dtype=dt)
    fit._alloc_pinned = unpinned
    try:
        key = (torch.float64, 12)
        fit._free_pinned.pop(key, None)
        lease = fit._pinned_buffer((4, 3), torch.float64)
        lease.tensor.fill_(2.0)
        arr = lease.numpy()
        ptr = lease.flat.data_ptr()
        view = arr[1:3, :2]
        del lease
        gc.collect()
        assert not fit._free_pinned.get(key) and arr.sum() == 24.0
        del arr
        gc.collect()
        assert not fit._free_pinned.get(key) and view.sum() == 8.0
        del view
        gc.collect()
        assert len(fit._free_pinned[key]) == 1
        again = fit._pinned_buffer((4, 3), torch.float64)
        assert again.flat.data_ptr() == ptr
        h = fit._HostResult(0, 5, 1, True)
        assert [a.shape for a in h.arrays()] == [(0, 5), (0, 5, 5), (0,), (0, 1), (0,), (0,)]
    finally:
        fit._alloc_pinned = real
