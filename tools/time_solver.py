"""Time the per-system solver kernels on a fixed batch: python tools/time_solver.py [nsys] [n]
(device ms per kernel kind from the library's own CUDA-event counters)."""
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
from volumetricinterp_b200 import _native, fit

S = int(sys.argv[1]) if len(sys.argv) > 1 else 28416
n = int(sys.argv[2]) if len(sys.argv) > 2 else 144
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
nrec = 64
Gs = []
for _ in range(nrec):
    M = rng.standard_normal((n, n)) * 10.0 ** rng.uniform(-6, 0, n)      # graded spectrum, like the fits
    Gs.append(M @ M.T)
G = torch.from_numpy(np.array(Gs)).to(dev)
y = torch.from_numpy(rng.standard_normal((nrec, n))).to(dev)
rec = torch.from_numpy((np.arange(S) % nrec).astype(np.int32)).to(dev)
regs = torch.from_numpy(np.eye(n)[None]).to(dev)
lam = torch.from_numpy(10.0 ** rng.uniform(-12, -2, (S, 1))).to(dev)
Cf = torch.empty((S, n), dtype=torch.float64, device=dev)
rank = torch.zeros((S,), dtype=torch.int32, device=dev)
status = torch.zeros((S,), dtype=torch.int32, device=dev)
ws = fit._workspace(dev, S, 1, n, 1, S)
lib = _native.lib()
lib.vi_profile_enable(1)
for it in range(2):
    lib.vi_profile_reset()
    _native.check(lib.vi_solve_batched(G.data_ptr(), y.data_ptr(), rec.data_ptr(), regs.data_ptr(), lam.data_ptr(), S, n, 1,
                                       2.220446049250313e-16, Cf.data_ptr(), rank.data_ptr(), status.data_ptr(),
                                       ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
ms, cnt = _native.profile_read()
print({k: round(v, 2) for k, v in ms.items() if cnt[k]}, "us/system tridiag(+chase):",
      round(1e3 * (ms["tridiag"] + ms.get("chase", 0.0)) / S, 3), "ok:", int((status == 0).sum().item()),
      "rank mean", float(rank.double().mean().item()))
