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
        self.hull_eq = np.ascontiguousarray(ConvexHull(self.hull_vert).equations)
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
        """estimate.py:180-221: nearest record within timetol, or linear interpolation in time."""
        t0 = (time - dt.datetime.utcfromtimestamp(0)).total_seconds()
        mt = np.mean(self.time, axis=1)
        try:
            if self.timeinterp:
                i = np.argwhere((t0 >= mt[:-1]) & (t0 < mt[1:])).flatten()[0]
                T = (t0 - mt[i]) / (mt[i + 1] - mt[i])
                C = (1 - T) * self.Coeffs[i, :] + T * self.Coeffs[i + 1, :]
                dC = None
                if self.Covariance is not None:
                    dC = (1 - T) * self.Covariance[i, :, :] + T * self.Covariance[i + 1, :, :]
            else:
                i = np.argmin(np.abs(mt - t0))
                if np.abs(mt[i] - t0) > self.timetol:
                    raise IndexError
                C = self.Coeffs[i]
                dC = self.Covariance[i] if self.Covariance is not None else None
        except IndexError:
            raise ValueError('Requested time out of range of data file.')
        return C, dC

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
