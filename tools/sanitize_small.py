"""Small end-to-end pass over every round-2 kernel at small and odd sizes (orders at the kernel boundaries, masked gates,
ragged tiles).  Written for `compute-sanitizer --tool memcheck`; the sanitizer is closed on this GPU pool, so the pass
was run plain (it completes; results are checked by the GPU tests)."""
import io
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import load_golden, product_model
from volumetricinterp_b200 import _native, fit

dev = torch.device("cuda", 0)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
rng = np.random.default_rng(0)
# fit at N = 27 and N = 144 (k_band + k_band_tail + k_chase + QL + k_replay_wave + chi2 + covariance)
for name in ("mid27", "c1_144"):
    g = load_golden(name)
    res = fit.fit_records(product_model(g), g["lat"], g["lon"], g["alt"], g["value"], g["error"], g["regs"], "chi2",
                          device=dev, want_cov=True, want_trace=True)
    print(name, "status", res.status.tolist(), "rank", res.rank.tolist())
# solver at orders around the kernel boundaries (168 / 169: k_band vs k_band_big), normal equations blocked kernel
for n in (9, 143, 168, 169, 200):
    M = rng.standard_normal((n, n))
    G = t((M @ M.T + n * np.eye(n))[None])
    y = t(rng.standard_normal((1, n)))
    S = 5
    rec = torch.zeros(S, dtype=torch.int32, device=dev)
    regs = t(np.eye(n)[None])
    lam = t(10.0 ** rng.uniform(-6, -2, (S, 1)))
    Cf = torch.empty((S, n), dtype=torch.float64, device=dev)
    rank = torch.zeros(S, dtype=torch.int32, device=dev)
    status = torch.zeros(S, dtype=torch.int32, device=dev)
    ws = fit._workspace(dev, S, 1, n, 1, S)
    _native.check(_native.lib().vi_solve_batched(G.data_ptr(), y.data_ptr(), rec.data_ptr(), regs.data_ptr(), lam.data_ptr(), S, n, 1,
                                                 2.220446049250313e-16, Cf.data_ptr(), rank.data_ptr(), status.data_ptr(),
                                                 ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    print("solve n", n, "rank", rank.tolist(), "status", status.tolist())
for N, P in ((161, 100), (300, 257)):
    A = t(rng.standard_normal((P, N)))
    v = rng.uniform(1e10, 1e12, (2, P)); v[0, ::5] = np.nan
    Gf, yf, *_ = fit.normal_equations_device(A, t(v), t(rng.uniform(1e9, 1e11, (2, P))), None, _native.NE_FAST)
    torch.cuda.synchronize()
    print("normal equations N", N, "finite", bool(torch.isfinite(Gf).all().item()))
# Estimate, many records (hull compaction + rows + GEMM + fill), odd sizes
g = load_golden("lo12")
from volumetricinterp_b200 import Estimate
import datetime as dt
model = product_model(g)
from scipy.spatial import ConvexHull
from volumetricinterp_b200.estimate import hull_halfspaces
eq = t(hull_halfspaces(g["hull_vert"]))
npts, R = 1001, 37
la = t(rng.uniform(g["q_lat"].min(), g["q_lat"].max(), npts)); lo = t(rng.uniform(g["q_lon"].min(), g["q_lon"].max(), npts))
al = t(rng.uniform(g["q_alt"].min(), g["q_alt"].max(), npts))
out = torch.empty((R, npts), dtype=torch.float64, device=dev)
model.estimate_device(la, lo, al, t(rng.standard_normal((R, model.nbasis))), eq, out)
torch.cuda.synchronize()
print("estimate many: inside fraction", float(torch.isfinite(out[0]).double().mean().item()))
print("sanitize pass complete")
