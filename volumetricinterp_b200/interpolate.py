"""`Interpolate` — drop-in for the reference class of the same name
(reference volumetricinterp/interpolate.py:16-708) with the record loop running on
the GPU.  Same constructor, config keys, attributes (`Coeffs`, `Covariance`,
`chi_sq`, `time`, `hull_vert`) and methods (`calc_coeffs`, `eval_C`,
`find_reg_param`, `compute_hull`, `read_datafile`, `saveh5`).
"""
import configparser
import datetime as dt
import importlib
import os

import numpy as np
from scipy.spatial import ConvexHull

from . import _native, fit as _fit
from .geo import geodetic2ecef


class Interpolate(object):

    def __init__(self, config_file):
        self.configfile = config_file
        self.read_config(self.configfile)
        # model plug-in protocol of the reference (interpolate.py:61-62)
        m = importlib.import_module('.models.' + self.model_name, package=__package__)
        with open(self.configfile) as f:
            self.model = m.Model(f)
        self.ne_mode = _native.NE_FAST
        self.calc_covariance = True

    # attribute <- (section, key, converter): the reference's configuration schema (interpolate.py:64-88)
    _CONFIG_SCHEMA = (
        ('param', 'DEFAULT', 'PARAM', str),
        ('filename', 'DEFAULT', 'FILENAME', str),
        ('outputfilename', 'DEFAULT', 'OUTPUTFILENAME', str),
        ('regularization_list', 'DEFAULT', 'REGULARIZATION_LIST', lambda v: [x for x in v.split(',') if x]),
        ('reg_method', 'DEFAULT', 'REGULARIZATION_METHOD', str),
        ('errlim', 'DEFAULT', 'ERRLIM', lambda v: [float(x) for x in v.split(',')]),
        ('chi2lim', 'DEFAULT', 'CHI2LIM', lambda v: [float(x) for x in v.split(',')]),
        ('goodfitcode', 'DEFAULT', 'GOODFITCODE', lambda v: [int(x) for x in v.split(',')]),
        ('model_name', 'MODEL', 'NAME', str),
    )

    def read_config(self, config_file):
        """Read the keys listed in _CONFIG_SCHEMA into attributes of the same names the reference uses."""
        parser = configparser.ConfigParser()
        with open(config_file) as f:
            parser.read_file(f)
        for attr, section, key, conv in self._CONFIG_SCHEMA:
            setattr(self, attr, conv(parser.get(section, key)))

    # ------------------------------------------------------------------ fit
    def eval_reg_matrices(self):
        """interpolate.py:486-493 (KeyError for a regulariser the model lacks)."""
        reg_matricies = {}
        for reg in self.regularization_list:
            try:
                reg_matricies[reg] = self.model.eval_reg_matricies[reg]()
            except KeyError as e:
                print('WARNING: The model {} does not support {} regularization!'.format(self.model_name, reg))
                raise e
        return reg_matricies

    def calc_coeffs(self, starttime=None, endtime=None, reg_matricies=None):
        """Fit every record of the file (interpolate.py:472-579) in one batched GPU pass."""
        if reg_matricies is None:
            print('Evaluating Regularization matricies.  This may take a few minutes.')
            reg_matricies = self.eval_reg_matrices()
        utime, lat, lon, alt, value, error = self.read_datafile(self.filename)
        self.compute_hull(lat, lon, alt)
        if starttime and endtime:
            t0 = (starttime - dt.datetime.utcfromtimestamp(0)).total_seconds()
            t1 = (endtime - dt.datetime.utcfromtimestamp(0)).total_seconds()
            idx = np.argwhere((utime[:, 0] >= t0) & (utime[:, 1] <= t1)).flatten()
            utime, value, error = utime[idx, :], value[idx], error[idx]
        if self.reg_method not in ('chi2', 'gcv') and self.regularization_list:
            # 'manual' and 'prompt' are broken in the reference itself (7-argument signatures called with 5,
            # interpolate.py:141 vs :353,:383)
            raise ValueError('REGULARIZATION_METHOD {} is not available (chi2, gcv)'.format(self.reg_method))
        res = _fit.fit_records(self.model, lat, lon, alt, value, error,
                               [reg_matricies[r] for r in self.regularization_list], self.reg_method,
                               ne_mode=self.ne_mode, want_cov=self.calc_covariance)
        self.time = utime
        self.Coeffs = res.Coeffs
        self.Covariance = res.Covariance
        self.chi_sq = res.chi_sq
        self.reg_params = res.reg_params
        self.fit_status = res.status
        self.fit_rank = res.rank
        return res

    # -------------------------------------------------- single-record operator seam
    def _one_record(self, A, b, W):
        import torch
        dev = torch.device('cuda', torch.cuda.current_device())
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        Ad = t(A)
        return Ad, Ad.t().contiguous(), t(np.asarray(b)[None, :]), t(np.asarray(W)[None, :])

    def eval_C(self, A, b, W, reg_matrices, reg_params, calccov=False):
        """interpolate.py:432-469 for one record (A (P,N), b (P,), W (P,)): C, or (C, dC) with calccov."""
        import torch
        Ad, At, bd, Wd = self._one_record(A, b, W)
        G, y, _, npts, Wm, bm = _fit.normal_equations_device(Ad, bd, None, Wd, self.ne_mode)
        names = [r for r in self.regularization_list if r in reg_params]
        N = Ad.shape[1]
        nreg = len(names)
        dev = Ad.device
        regs = torch.from_numpy(np.stack([reg_matrices[r] for r in names])).to(dev) if nreg else None
        lam = torch.tensor([[float(reg_params[r]) for r in names]], dtype=torch.float64, device=dev) if nreg else None
        Cf = torch.empty((1, N), dtype=torch.float64, device=dev)
        dC = torch.empty((1, N, N), dtype=torch.float64, device=dev) if calccov else None
        rank = torch.zeros((1,), dtype=torch.int32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        ws = _fit._workspace(dev, 1, Ad.shape[0], N, nreg, 32)
        _native.check(_native.lib().vi_solve_cov_batched(
            G.data_ptr(), y.data_ptr(), None, regs.data_ptr() if nreg else None, lam.data_ptr() if nreg else None,
            1, N, nreg, np.finfo(float).eps, Cf.data_ptr(), dC.data_ptr() if calccov else None, rank.data_ptr(),
            status.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(dev).cuda_stream))
        if int(status.item()) == _native.ST_NONFINITE:
            raise ValueError('array must not contain infs or NaNs')
        if calccov:
            return Cf[0].cpu().numpy(), dC[0].cpu().numpy()
        return Cf[0].cpu().numpy()

    def find_reg_param(self, A, b, W, reg_matrices, method=None):
        """interpolate.py:97-147 for one record: {name: lambda or NaN}."""
        import torch
        if method not in (None, 'chi2', 'gcv'):
            raise ValueError('only the chi2 and gcv methods are available')
        meth = _native.METHOD_GCV if method == 'gcv' else _native.METHOD_CHI2
        Ad, At, bd, Wd = self._one_record(A, b, W)
        G, y, _, npts, Wm, bm = _fit.normal_equations_device(Ad, bd, None, Wd, self.ne_mode)
        out = {}
        for name in self.regularization_list:
            regs = torch.from_numpy(np.asarray(reg_matrices[name], dtype=np.float64)[None]).to(Ad.device)
            _, _, _, lam, _, status, _ = _fit.fit_batch_device(At, Wm, bm, G, y, npts, regs, meth, A=Ad)
            out[name] = float(lam[0, 0].item())
        return out

    # ------------------------------------------------------------------ geometry / io
    def compute_hull(self, lat, lon, alt):
        """interpolate.py:409-426: hull vertices of the gate cloud in ECEF."""
        x, y, z = geodetic2ecef(lat, lon, alt)
        R0 = np.array([x, y, z]).T
        chull = ConvexHull(R0)
        self.hull_vert = R0[chull.vertices]

    def quality_filter(self, lat, lon, alt, val, err, chi2, fitcode):
        """The in-memory half of read_datafile (interpolate.py:635-667)."""
        altitude, latitude, longitude = alt.flatten(), lat.flatten(), lon.flatten()
        chi2 = chi2.reshape(chi2.shape[0], -1)
        fitcode = fitcode.reshape(fitcode.shape[0], -1)
        value = np.array(val.reshape(val.shape[0], -1), dtype=float)
        error = np.array(err.reshape(err.shape[0], -1), dtype=float)
        # AMISR chi2 offset quirk (interpolate.py:645-646)
        if np.nanmedian(chi2) > 100.:
            chi2 = chi2 - 369.
        with np.errstate(invalid='ignore'):
            good = ((error > self.errlim[0]) & (error < self.errlim[1]) & (chi2 > self.chi2lim[0])
                    & (chi2 < self.chi2lim[1]) & np.isin(fitcode, self.goodfitcode))
        value[~good] = np.nan
        error[~good] = np.nan
        keep = np.isfinite(altitude)
        return latitude[keep], longitude[keep], altitude[keep], value[:, keep], error[:, keep]

    def read_datafile(self, filename):
        """interpolate.py:582-667: AMISR fitted HDF5 -> (utime, lat, lon, alt, value, error)."""
        from . import h5lite
        with h5lite.File(filename) as h5:
            utime = h5['/Time/UnixTime']
            alt = h5['/Geomag/Altitude']
            lat = h5['/Geomag/Latitude']
            lon = h5['/Geomag/Longitude']
            c2 = h5['/FittedParams/FitInfo/chi2']
            fc = h5['/FittedParams/FitInfo/fitcode']
            if self.param == 'dens':
                val = h5['/FittedParams/Ne']
                err = h5['/FittedParams/dNe']
            else:
                # <kind>_<ion>: Fits/Errors[..., ion_index, kind_index] (interpolate.py:615-632)
                kind, ion = self.param.split('_')
                m = {'O': 16, 'O2': 32, 'NO': 30, 'N2': 28, 'N': 14}[ion]
                i = {'frac': 0, 'temp': 1, 'colfreq': 2}[kind]
                mass = h5['/FittedParams/IonMass']
                try:
                    mi = int(np.where(mass == m)[0][0])
                except IndexError:
                    mi = -1            # reference falls back to the last species (interpolate.py:626-629)
                val = h5['/FittedParams/Fits'][:, :, :, mi, i]
                err = h5['/FittedParams/Errors'][:, :, :, mi, i]
        lat, lon, alt, value, error = self.quality_filter(lat, lon, alt, val, err, c2, fc)
        return utime, lat, lon, alt, value, error

    def saveh5(self):
        """interpolate.py:671-708: coefficient file, same layout."""
        from . import h5lite
        with open(self.configfile) as f:
            contents = f.read()
        path = os.path.dirname(os.path.abspath(self.configfile))
        name = os.path.basename(self.configfile)
        R, N = self.Coeffs.shape
        cov = self.Covariance if self.Covariance is not None else np.full((R, N, N), np.nan)
        with h5lite.Writer(self.outputfilename) as h5:
            h5.array('/UnixTime', self.time)
            h5.group('/Coeffs', title='Dataset')
            h5.group('/FitParams', title='Dataset')
            h5.group('/RawData', title='Dataset')
            h5.array('/Coeffs/C', self.Coeffs)
            h5.array('/Coeffs/dC', cov)
            h5.strings('/FitParams/reglist', self.regularization_list)
            h5.string('/FitParams/regmethod', self.reg_method.encode('utf-8'))
            h5.array('/FitParams/chi2', self.chi_sq)
            h5.array('/FitParams/hull_vert', self.hull_vert)
            h5.string('/RawData/filename', self.filename.encode('utf-8'))
            h5.group('/ConfigFile')
            h5.string('/ConfigFile/Name', name.encode('utf-8'))
            h5.string('/ConfigFile/Path', path.encode('utf-8'))
            h5.string('/ConfigFile/Contents', contents.encode('utf-8'))
