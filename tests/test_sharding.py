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
