"""`radbasfun` model plug-in: Gaussian radial basis functions on a geodetic grid.

Host-side mirror of the reference plug-in (reference
volumetricinterp/models/radbasfun.py:52-112): `Model(config_file)` with `nbasis`,
`eval_reg_matricies` (empty: the reference model offers no regulariser, :62) and
`basis(gdlat, gdlon, gdalt) -> shape+(N,)`.  The basis is evaluated by the CUDA
kernel `vi_basis_radbasfun` (csrc/basis.cu).
"""
import configparser

import numpy as np

from .. import _native
from ..geo import geodetic2ecef


class Model(object):
    name = "radbasfun"

    def __init__(self, config_file):
        cfg = configparser.ConfigParser()
        cfg.read_file(config_file)
        # keys of reference radbasfun.py:70-78
        self.latcp = cfg.getfloat('MODEL', 'LATCP')
        self.loncp = cfg.getfloat('MODEL', 'LONCP')
        self.eps = cfg.getfloat('MODEL', 'EPS')
        self.latrange = [float(i) for i in cfg.get('MODEL', 'LATRANGE').split(',')]
        self.lonrange = [float(i) for i in cfg.get('MODEL', 'LONRANGE').split(',')]
        self.altrange = [float(i) for i in cfg.get('MODEL', 'ALTRANGE').split(',')]
        self.numgridpnt = cfg.getint('MODEL', 'NUMGRIDPNT')
        g = self.numgridpnt
        # centres: meshgrid (default 'xy' indexing) flattened C-order, km -> m (radbasfun.py:55-59)
        lat, lon, alt = np.meshgrid(np.linspace(self.latrange[0], self.latrange[1], g),
                                    np.linspace(self.lonrange[0], self.lonrange[1], g),
                                    np.linspace(self.altrange[0], self.altrange[1], g) * 1000.)
        x, y, z = geodetic2ecef(lat.flatten(), lon.flatten(), alt.flatten())
        self.centers = np.ascontiguousarray(np.array([x, y, z]).T)
        self.nbasis = self.centers.shape[0]
        self.eval_reg_matricies = {}
        self._centers_dev = {}

    def centers_device(self, device):
        import torch
        key = str(device)
        if key not in self._centers_dev:
            self._centers_dev[key] = torch.from_numpy(self.centers).to(device)
        return self._centers_dev[key]

    def basis_device(self, lat, lon, alt, out=None, out_t=None, stream=None):
        """1-D float64 CUDA tensors in, A (npts, N) row-major out; optionally out_t (N, npts)."""
        import torch
        npts = lat.numel()
        if out is None and out_t is None:
            out = torch.empty((npts, self.nbasis), dtype=torch.float64, device=lat.device)
        s = stream if stream is not None else torch.cuda.current_stream(lat.device).cuda_stream
        cen = self.centers_device(lat.device)
        _native.check(_native.lib().vi_basis_radbasfun(
            lat.data_ptr(), lon.data_ptr(), alt.data_ptr(), npts, cen.data_ptr(), self.nbasis, self.eps,
            out.data_ptr() if out is not None else None, out_t.data_ptr() if out_t is not None else None, s))
        return out

    def estimate_device(self, lat, lon, alt, C, hull_eq, out, stream=None, ws_slot=0):
        """out[r, p] = basis(p) . C[r], NaN outside the hull (estimate.py:113-121); 16 or more records go through the
        compaction + FP64 tensor-core GEMM path (vi_estimate_radbasfun_many)."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(lat.device).cuda_stream
        cen = self.centers_device(lat.device)
        F = 0 if hull_eq is None else hull_eq.shape[0]
        npts, R = lat.numel(), C.shape[0]
        if R >= 16 and npts > 0 and self.nbasis <= 144:      # the GEMM tile holds K = nbasis columns in shared memory
            ws = _native.estimate_workspace(lat.device, npts, self.nbasis, R, ws_slot)
            _native.check(_native.lib().vi_estimate_radbasfun_many(
                lat.data_ptr(), lon.data_ptr(), alt.data_ptr(), npts, cen.data_ptr(), self.nbasis, self.eps,
                C.data_ptr(), R, hull_eq.data_ptr() if F else None, F, out.data_ptr(), ws.data_ptr(), ws.numel(), s))
            return out
        _native.check(_native.lib().vi_estimate_radbasfun(
            lat.data_ptr(), lon.data_ptr(), alt.data_ptr(), npts, cen.data_ptr(), self.nbasis, self.eps,
            C.data_ptr(), R, hull_eq.data_ptr() if F else None, F, out.data_ptr(), s))
        return out

    def basis(self, gdlat, gdlon, gdalt):
        """Drop-in for reference radbasfun.py:83-112 (numpy in, numpy out)."""
        import torch
        gdlat, gdlon, gdalt = (np.asarray(a, dtype=np.float64) for a in (gdlat, gdlon, gdalt))
        dev = torch.device('cuda', torch.cuda.current_device())
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel())).to(dev)
        A = self.basis_device(to(gdlat), to(gdlon), to(gdalt))
        return A.cpu().numpy().reshape(gdlat.shape + (self.nbasis,))
