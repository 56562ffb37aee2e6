"""`Estimate` — drop-in for the reference class (reference
volumetricinterp/estimate.py:13-221): `Estimate(file)(time, gdlat, gdlon, gdalt)`.

The basis evaluation, the dot product with the coefficients and the in-hull mask
run in one CUDA kernel (`vi_estimate_*`, csrc/basis.cu).  The reference rebuilds
a Qhull per query point (estimate.py:167-177); the same decision is taken here
from the facet half-spaces of ConvexHull(hull_vert), computed once on the host.
"""
import configparser
import datetime as dt
import importlib
import io

import numpy as np
from scipy.spatial import ConvexHull


def hull_halfspaces(hull_vert):
    """Facet half-spaces [n | d] of ConvexHull(hull_vert) for the in-kernel test n.x + d <= 0, with the
    distance tolerance of the reference's decision folded into d.  The reference asks Qhull whether
    hull_vert + {p} has the same vertices as hull_vert (estimate.py:167-177); Qhull only makes p a vertex when it
    lies farther outside a facet than its roundoff allowance (qh_detroundoff: DISTround =
    eps (1.01 dim sqrt(dim) + 1) max|coord|, a point counts as outside beyond about 6 DISTround when coplanar
    facets are merged, the 3-d default), so gates ON the hull -- its own vertices, points on facets -- are inside.
    A bare `<= 0` decides those by rounding noise of the ECEF coordinates (~1e-9 m at 6.4e6 m)."""
    hv = np.asarray(hull_vert, dtype=np.float64)
    eq = np.array(ConvexHull(hv).equations, dtype=np.float64, order='C')
    dist_round = np.finfo(np.float64).eps * (1.01 * 3.0 * np.sqrt(3.0) + 1.0) * np.abs(hv).max()
    eq[:, 3] -= 6.0 * dist_round
    return eq


class Estimate(object):

    def __init__(self, param_file, timetol=60., timeinterp=False):
        self.timetol = timetol
        self.timeinterp = timeinterp
        self.loadh5(filename=param_file)
        self._init_model(self.config_text)

    @classmethod
    def from_arrays(cls, config_text, time, Coeffs, hull_vert, Covariance=None, timetol=60., timeinterp=False):
        """Build an Estimate without a coefficient file (tests, pipelines)."""
        self = cls.__new__(cls)
        self.timetol, self.timeinterp = timetol, timeinterp
        self.time, self.Coeffs, self.Covariance = np.asarray(time), np.asarray(Coeffs), Covariance
        self.hull_vert = np.asarray(hull_vert)
        self.config_text = config_text
        self._init_model(config_text)
        return self

    def _init_model(self, config_text):
        config = configparser.ConfigParser()
        config.read_file(io.StringIO(config_text))
        self.model_name = config.get('MODEL', 'NAME')
        m = importlib.import_module('.models.' + self.model_name, package=__package__)
        self.model = m.Model(io.StringIO(config_text))
        self.hull_eq = hull_halfspaces(self.hull_vert)
        self._dev = {}

    def loadh5(self, filename):
        """estimate.py:53-70."""
        from . import h5lite
        with h5lite.File(filename) as h5:
            self.Coeffs = h5['/Coeffs/C']
            self.Covariance = h5['/Coeffs/dC']
            self.time = h5['/UnixTime']
            self.hull_vert = h5['/FitParams/hull_vert']
            txt = h5['/ConfigFile/Contents']
        self.config_text = txt.decode('utf-8') if isinstance(txt, (bytes, np.bytes_)) else str(txt)

    def get_C(self, time):
        """Coefficients (and covariance) in force at `time` -- the selection rule of estimate.py:180-221: the record
        whose mid-time is closest, provided it is within `timetol` seconds; with `timeinterp` the two records whose
        mid-times bracket `time`, blended linearly.  ValueError when the file does not cover `time`."""
        t = (time - dt.datetime.utcfromtimestamp(0)).total_seconds()
        mid = np.asarray(self.time, dtype=float).mean(axis=1)
        cov = self.Covariance
        if self.timeinterp:
            # left neighbour: last record with mid <= t (records are in file order, as the reference assumes)
            hit = np.flatnonzero((mid[:-1] <= t) & (t < mid[1:]))
            if hit.size == 0:
                raise ValueError('Requested time out of range of data file.')
            j = int(hit[0])
            frac = (t - mid[j]) / (mid[j + 1] - mid[j])
            blend = lambda a: (1 - frac) * a[j] + frac * a[j + 1]
            return blend(self.Coeffs), (blend(cov) if cov is not None else None)
        if mid.size == 0:
            raise ValueError('Requested time out of range of data file.')
        gap = np.abs(mid - t)
        j = int(np.argmin(gap))
        if gap[j] > self.timetol:
            raise ValueError('Requested time out of range of data file.')
        return self.Coeffs[j], (cov[j] if cov is not None else None)

    def _device_consts(self, dev):
        import torch
        key = str(dev)
        if key not in self._dev:
            self._dev[key] = torch.from_numpy(self.hull_eq).to(dev)
        return self._dev[key]

    def evaluate_device(self, C, lat, lon, alt, check_hull=True, out=None):
        """C (Rsel,N), lat/lon/alt (npts,) CUDA float64 tensors -> out (Rsel, npts)."""
        import torch
        dev = lat.device
        if out is None:
            out = torch.empty((C.shape[0], lat.numel()), dtype=torch.float64, device=dev)
        eq = self._device_consts(dev) if check_hull else None
        self.model.estimate_device(lat, lon, alt, C, eq, out)
        return out

    def __call__(self, time, gdlat, gdlon, gdalt, calcgrad=False, calcerr=False, check_hull=True):
        """estimate.py:75-123.  calcgrad / calcerr are accepted and ignored, as in the reference."""
        import torch
        C, _ = self.get_C(time)
        gdlat, gdlon, gdalt = (np.asarray(a, dtype=np.float64) for a in (gdlat, gdlon, gdalt))
        dev = torch.device('cuda', torch.cuda.current_device())
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel())).to(dev, non_blocking=True)
        Cd = torch.from_numpy(np.ascontiguousarray(C, dtype=np.float64)[None, :]).to(dev)
        out = self.evaluate_device(Cd, to(gdlat), to(gdlon), to(gdalt), check_hull)
        return out[0].cpu().numpy().reshape(gdlat.shape)
