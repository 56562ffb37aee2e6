"""Diagnostic (GPU box): bench-shaped workload, GPU stages vs the oracle on a handful of records."""
import io, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import bench
import ref_port as rp
from volumetricinterp_b200 import _native, fit
from volumetricinterp_b200.models import sphharmlag

dev = torch.device("cuda", 0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 8
model = sphharmlag.Model(io.StringIO(bench.config_text()))
lat, lon, alt = bench.make_geometry(seed=100)
f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
la, lo, al = f64(lat), f64(lon), f64(alt)
P, N = len(lat), model.nbasis
A = torch.empty((P, N), dtype=torch.float64, device=dev); At = torch.empty((N, P), dtype=torch.float64, device=dev)
model.basis_device(la, lo, al, out=A, out_t=At)
Ah = A.cpu().numpy()
value, error = bench.make_workload(Ah, R, seed=200)
Om = bench.curvature_matrix()
v, e = f64(value), f64(error)
with np.errstate(invalid="ignore"):
    W = error ** -2
Gf, yf, _, npts, Wm, bm = fit.normal_equations_device(A, v, e, None, _native.NE_FAST)
Gs, ys, *_ = fit.normal_equations_device(A, v, e, f64(W), _native.NE_STRICT)
Gf, yf, Gs, ys = (t.cpu().numpy() for t in (Gf, yf, Gs, ys))
for r in range(min(R, 4)):
    ok = np.isfinite(value[r])
    G, y = rp.normal_equations(Ah[ok], W[r][ok], value[r][ok])
    print("rec", r, "strict==einsum", np.array_equal(Gs[r], G), np.array_equal(ys[r], y),
          "fast rel", np.abs(Gf[r] - G).max() / np.abs(G).max(), np.abs(yf[r] - y).max() / np.abs(y).max())
res = fit.fit_records(model, lat, lon, alt, value, error, [Om], "chi2", device=dev)
print("status", res.status, "lam", res.reg_params.ravel(), "chi2/n", res.chi_sq / np.isfinite(value).sum(1), "rank", res.rank)
# table of the first record from the oracle vs GPU single solves
import ctypes as C
def gpu_solve(G, y, lam):
    S = len(lam)
    Cf = torch.empty((S, N), dtype=torch.float64, device=dev); rank = torch.zeros((S,), dtype=torch.int32, device=dev)
    st = torch.zeros((S,), dtype=torch.int32, device=dev)
    ws = fit._workspace(dev, S, 1, N, 1, max(32, S))
    Gd, yd, rd, ld = f64(np.repeat(G[None], S, 0)), f64(np.repeat(y[None], S, 0)), f64(Om[None]), f64(np.array(lam)[:, None])
    _native.check(_native.lib().vi_solve_batched(Gd.data_ptr(), yd.data_ptr(), None, rd.data_ptr(), ld.data_ptr(), S, N, 1,
        np.finfo(float).eps, Cf.data_ptr(), rank.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    return Cf.cpu().numpy(), rank.cpu().numpy(), st.cpu().numpy()
import scipy.linalg
for r in range(min(R, 3)):
    ok = np.isfinite(value[r]); Ar = Ah[ok]; b = value[r][ok]; Wr = W[r][ok]
    G, y = rp.normal_equations(Ar, Wr, b)
    alphas = [0.0, -5.0, -15.0, -20.0, -22.0, -24.0, -26.0, -30.0, -40.0]
    Cg, rk, st = gpu_solve(G, y, [10.0 ** a for a in alphas])
    for i, a in enumerate(alphas):
        X = G + 10.0 ** a * Om
        Cl = scipy.linalg.lstsq(X, y)[0]
        s = np.linalg.svd(X, compute_uv=False)
        chi = lambda Cx: np.sum((Ar @ Cx - b) ** 2 * Wr)
        print(f"rec {r} alpha {a}: gpu rank {rk[i]} st {st[i]} ref rank {(s > np.finfo(float).eps * s[0]).sum()} chi2/n gpu {chi(Cg[i]) / ok.sum():.6f} ref {chi(Cl) / ok.sum():.6f}")
