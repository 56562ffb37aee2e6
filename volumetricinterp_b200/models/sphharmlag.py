"""`sphharmlag` model plug-in: Laguerre x spherical-cap-harmonic basis.

Host-side mirror of the reference plug-in protocol (reference
volumetricinterp/models/sphharmlag.py:11-15, 57-62): `Model(config_file)` with
`nbasis`, `eval_reg_matricies` and `basis(gdlat, gdlon, gdalt) -> shape+(N,)`.
The basis itself is evaluated by the CUDA kernel `vi_basis_sphharmlag`
(csrc/basis.cu); this class only parses the reference's [MODEL] keys, tabulates
the per-(l,|m|) constants with the same scipy calls the reference makes, and
owns the (host-side, config-only) regularisation matrices.
"""
import configparser
import ctypes as C

import numpy as np
import scipy.integrate
import scipy.special as sp

from .. import _native
from ..geo import geodetic2ecef

RE = 6371.2 * 1000.0


class Model(object):
    name = "sphharmlag"

    def __init__(self, config_file):
        cfg = configparser.ConfigParser()
        cfg.read_file(config_file)
        # keys of reference sphharmlag.py:70-75
        self.maxk = cfg.getint('MODEL', 'MAXK')
        self.maxl = cfg.getint('MODEL', 'MAXL')
        self.latcp = cfg.getfloat('MODEL', 'LATCP')
        self.loncp = cfg.getfloat('MODEL', 'LONCP')
        self.cap_lim = cfg.getfloat('MODEL', 'CAP_LIM') * np.pi / 180.
        self.max_z_int = float(cfg.get('MODEL', 'MAX_Z_INT'))
        self.nbasis = self.maxk * self.maxl**2
        self.eval_reg_matricies = {'curvature': self.eval_omega, '0thorder': self.eval_psi}
        self._params = None

    # ---- index bookkeeping (reference sphharmlag.py:79-115) -------------
    def basis_numbers(self, n):
        r = n % (self.maxl**2)
        l = np.floor(np.sqrt(r))
        return n // (self.maxl**2), l, r - l * (l + 1)

    def nu(self, n):
        _, l, _ = self.basis_numbers(n)
        return self._nu_l(l)

    def _nu_l(self, l):
        return (2 * l + 0.5) * np.pi / (2 * self.cap_lim) - 0.5

    def Kvm(self, v, m):
        k = np.sqrt((2 * v + 1) / (4 * np.pi) * sp.gamma(float(v - m + 1)) / sp.gamma(float(v + m + 1)))
        return k * np.sqrt(2) if m != 0 else k

    def Az(self, v, m, phi):
        am = abs(m)
        return self.Kvm(v, am) * (np.sin(am * phi) if m < 0 else np.cos(am * phi))

    # ---- device parameter block ---------------------------------------
    def params(self):
        """`vi_shl_params` for the CUDA kernels (csrc/vi_math.h)."""
        if self._params is None:
            x0, y0, z0 = geodetic2ecef(self.latcp, self.loncp, 0.)
            theta0 = np.arccos(z0 / np.sqrt(x0**2 + y0**2 + z0**2))
            phi0 = np.arctan2(y0, x0)
            L = self.maxl
            nu = [self._nu_l(float(l)) for l in range(L)]
            kvm = [[self.Kvm(nu[l], m) for m in range(l + 1)] for l in range(L)]
            with np.errstate(over='ignore'):
                g1 = [[sp.gamma(nu[l] - m + 1) for m in range(l + 1)] for l in range(L)]
                g2 = [[sp.gamma(nu[l] + m + 1) for m in range(l + 1)] for l in range(L)]
            self._params = _native.fill_shl_params(
                self.maxk, L, np.cos(theta0), np.sin(theta0),
                np.cos(phi0 + np.pi / 2.), np.sin(phi0 + np.pi / 2.), nu, kvm, g1, g2)
        return self._params

    # ---- basis ------------------------------------------------------------
    def basis_device(self, lat, lon, alt, out=None, out_t=None, stream=None):
        """lat/lon/alt: 1-D float64 CUDA tensors (deg, deg, m).  Returns A (npts, N)
        row-major on the device; optionally also fills out_t (N, npts)."""
        import torch
        npts = lat.numel()
        if out is None and out_t is None:
            out = torch.empty((npts, self.nbasis), dtype=torch.float64, device=lat.device)
        s = stream if stream is not None else torch.cuda.current_stream(lat.device).cuda_stream
        _native.check(_native.lib().vi_basis_sphharmlag(
            lat.data_ptr(), lon.data_ptr(), alt.data_ptr(), npts, C.byref(self.params()),
            out.data_ptr() if out is not None else None, out_t.data_ptr() if out_t is not None else None, s))
        return out

    def estimate_device(self, lat, lon, alt, Cf, hull_eq, out, stream=None, ws_slot=0):
        """out[r, p] = basis(p) . Cf[r], NaN outside the hull (estimate.py:113-121).  With 16 or more records the
        in-hull points are compacted and the contraction runs as an FP64 tensor-core GEMM (vi_estimate_*_many)."""
        import torch
        s = stream if stream is not None else torch.cuda.current_stream(lat.device).cuda_stream
        F = 0 if hull_eq is None else hull_eq.shape[0]
        npts, R = lat.numel(), Cf.shape[0]
        if R >= 16 and npts > 0 and self.nbasis <= 144:      # the GEMM tile holds K = nbasis columns in shared memory
            ws = _native.estimate_workspace(lat.device, npts, self.nbasis, R, ws_slot)
            _native.check(_native.lib().vi_estimate_sphharmlag_many(
                lat.data_ptr(), lon.data_ptr(), alt.data_ptr(), npts, C.byref(self.params()),
                Cf.data_ptr(), R, hull_eq.data_ptr() if F else None, F, out.data_ptr(), ws.data_ptr(), ws.numel(), s))
            return out
        _native.check(_native.lib().vi_estimate_sphharmlag(
            lat.data_ptr(), lon.data_ptr(), alt.data_ptr(), npts, C.byref(self.params()),
            Cf.data_ptr(), R, hull_eq.data_ptr() if F else None, F, out.data_ptr(), s))
        return out

    def basis(self, gdlat, gdlon, gdalt):
        """Drop-in for reference sphharmlag.py:118-145 (numpy in, numpy out)."""
        import torch
        gdlat, gdlon, gdalt = (np.asarray(a, dtype=np.float64) for a in (gdlat, gdlon, gdalt))
        dev = torch.device('cuda', torch.cuda.current_device())
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel())).to(dev)
        A = self.basis_device(to(gdlat), to(gdlon), to(gdalt))
        return A.cpu().numpy().reshape(gdlat.shape + (self.nbasis,))

    def grad_basis_device(self, lat, lon, alt, out=None, stream=None):
        """Gradient of the basis on the device: (npts, 3, N), components along z-hat, theta-hat, phi-hat."""
        import torch
        npts = lat.numel()
        if out is None:
            out = torch.empty((npts, 3, self.nbasis), dtype=torch.float64, device=lat.device)
        s = stream if stream is not None else torch.cuda.current_stream(lat.device).cuda_stream
        _native.check(_native.lib().vi_grad_basis_sphharmlag(
            lat.data_ptr(), lon.data_ptr(), alt.data_ptr(), npts, C.byref(self.params()), out.data_ptr(), s))
        return out

    def grad_basis(self, gdlat, gdlon, gdalt):
        """Drop-in for reference sphharmlag.py:148-184 (numpy in, numpy out; 1-D inputs as the reference's own
        transform_coord requires there).  Shape (npoints, 3, nbasis), as `np.array(Ag).T` in the reference."""
        import torch
        gdlat, gdlon, gdalt = (np.asarray(a, dtype=np.float64) for a in (gdlat, gdlon, gdalt))
        dev = torch.device('cuda', torch.cuda.current_device())
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a.ravel())).to(dev)
        return self.grad_basis_device(to(gdlat), to(gdlon), to(gdalt)).cpu().numpy()

    # ---- regularisation matrices (host, config-only) ---
    def _assemble(self, zfun, tfun):
        """N(N+1)/2 triple products of three 1-D QUADPACK integrals (reference
        sphharmlag.py:188-239).  The integrals are separable, so each distinct
        (ki,kj) / ordered (li,mi,lj,mj) integral is evaluated once with the same
        `quad` call and reused — bit-identical values, ~8x fewer quadratures."""
        N = self.nbasis
        zc, tc, pc = {}, {}, {}
        out = np.zeros((N, N))
        for ni in range(N):
            ki, li, mi = self.basis_numbers(ni)
            vi = self._nu_l(li)
            for nj in range(ni, N):
                kj, lj, mj = self.basis_numbers(nj)
                vj = self._nu_l(lj)
                kz = (ki, kj)
                if kz not in zc:
                    zc[kz] = scipy.integrate.quad(zfun(ki, kj), 0., self.max_z_int)[0]
                kt = (li, mi, lj, mj)
                if kt not in tc:
                    tc[kt] = scipy.integrate.quad(tfun(mi, vi, mj, vj), 0., self.cap_lim)[0]
                    pc[kt] = scipy.integrate.quad(
                        lambda p: self.Az(vi, mi, p) * self.Az(vj, mj, p), 0., 2 * np.pi)[0]
                out[ni, nj] = out[nj, ni] = zc[kz] * tc[kt] * pc[kt]
        return out

    def eval_omega(self):
        """curvature matrix (reference sphharmlag.py:188-212)."""
        def zfun(ki, kj):
            return lambda z: np.exp(-1 * z) * sp.eval_laguerre(ki, z) * sp.eval_laguerre(kj, z) / z**2

        def lap(m, v, t):
            c = np.cos(t)
            return (-1 * v * (v * c**2 + v + 1) * sp.lpmv(m, v, c) + v * (v + m) * c * sp.lpmv(m, v - 1, c)
                    + v * (v - m + 1) * c * sp.lpmv(m, v + 1, c))

        def tfun(mi, vi, mj, vj):
            return lambda t: 1 / np.sin(t)**3 * lap(mi, vi, t) * lap(mj, vj, t)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', scipy.integrate.IntegrationWarning)
            return self._assemble(zfun, tfun)

    def eval_psi(self):
        """0th-order matrix (reference sphharmlag.py:215-239)."""
        def zfun(ki, kj):
            return lambda z: np.exp(-1 * z) * sp.eval_laguerre(ki, z) * sp.eval_laguerre(kj, z) * z**2

        def tfun(mi, vi, mj, vj):
            return lambda t: sp.lpmv(mi, vi, np.cos(t)) * sp.lpmv(mj, vj, np.cos(t)) * np.sin(t)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore', scipy.integrate.IntegrationWarning)
            return self._assemble(zfun, tfun)
