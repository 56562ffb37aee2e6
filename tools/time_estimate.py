"""Time the many-record Estimate path on one tile of the C4 grid: python tools/time_estimate.py [records] [iters]
(random coefficients: the kernels do not care; tile from the middle of the 512^3 grid, ~1/3 of it inside the hull)."""
import io
import sys
import numpy as np
import torch

sys.path.insert(0, ".")
import bench
from scipy.spatial import ConvexHull
from volumetricinterp_b200 import _native
from volumetricinterp_b200.estimate import hull_halfspaces
from volumetricinterp_b200.geo import geodetic2ecef
from volumetricinterp_b200.models import sphharmlag

R = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
bench.NBEAMS, bench.NGATES = 51, 100
model = sphharmlag.Model(io.StringIO(bench.config_text()))
lat, lon, alt = bench.make_geometry(seed=100)
pts = np.array(geodetic2ecef(lat, lon, alt)).T
eq = torch.from_numpy(hull_halfspaces(pts[ConvexHull(pts).vertices])).to(dev)
G, tile = 512, 1 << 17
f64 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
gl, gn, ga = f64(np.linspace(lat.min(), lat.max(), G)), f64(np.linspace(lon.min(), lon.max(), G)), f64(np.linspace(100e3, 650e3, G))
p = torch.arange(G ** 3 // 2, G ** 3 // 2 + tile, device=dev, dtype=torch.int64)
la, lo, al = gl[p // (G * G)], gn[(p // G) % G], ga[p % G]
Cm = torch.from_numpy(np.random.default_rng(0).standard_normal((R, model.nbasis))).to(dev)
out = torch.empty((R, tile), dtype=torch.float64, device=dev)
lib = _native.lib()
for _ in range(3):
    model.estimate_device(la, lo, al, Cm, eq, out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib.vi_profile_reset()
lib.vi_profile_enable(1)
e0.record()
for _ in range(iters):
    model.estimate_device(la, lo, al, Cm, eq, out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
kms, kn = _native.profile_read()
lib.vi_profile_enable(0)
print({k: round(v / iters, 4) for k, v in kms.items() if kn[k]}, "ms per tile by kind")
fin = float(torch.isfinite(out[0]).double().mean().item())
print(f"records {R} tile {tile} inside {fin:.3f}: {ms:.3f} ms per tile, {tile * R / ms / 1e6:.1f} G(point,record)/s, "
      f"{2 * model.nbasis * tile * R * fin / ms / 1e9:.2f} TFLOP/s, {tile * R * 8 / ms / 1e6:.0f} GB/s written")
# spot check against the single-record kernel
ref = torch.empty((1, tile), dtype=torch.float64, device=dev)
model.estimate_device(la, lo, al, Cm[:1].contiguous(), eq, ref)
ok = torch.isfinite(ref[0])
print("mask identical:", bool((ok == torch.isfinite(out[0])).all().item()), "max rel diff:",
      float(((out[0][ok] - ref[0][ok]).abs().max() / ref[0][ok].abs().max()).item()))
