"""Device driver of the fit hot path: records in, coefficients out.

Python stays a thin host layer (device memory through torch, kernels through the
C ABI in _native.py).  One call fits all R records of a file:

    K1  design matrix A (P x N) and its transpose, once per file      (vi_basis_*)
    K2  masked weights + normal equations for a batch of records      (vi_normal_eq_batched)
    K3  chi^2 = nu search, final solve, chi^2                         (vi_fit_batched)

which replaces the reference's record loop, interpolate.py:511-574.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native

METHODS = {None: _native.METHOD_CHI2, 'chi2': _native.METHOD_CHI2, 'gcv': _native.METHOD_GCV}


@dataclass
class FitResult:
    Coeffs: object        # (R, N)
    Covariance: object    # (R, N, N) or None
    chi_sq: object        # (R,)
    reg_params: object    # (R, nreg)
    rank: object          # (R,) int32
    status: object        # (R,) int32, _native.ST_*
    nsolve: int = 0       # eigen-systems solved
    trace: object = None  # want_trace: dict(table (R*nreg, 102), nu, k_lo, kdone) of the chi2 search
    device_small: object = None   # to_host=True: the per-record results still on the device (for the gather)


def _cuda_stream(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


def normal_equations_device(A, value, error, weight=None, mode=_native.NE_FAST, want_masked=True):
    """A (P,N), value/error (R,P) CUDA float64 tensors -> G (R,N,N), y (R,N), sWbb (R), npts (R), Wm, bm."""
    import torch
    R, P = value.shape
    N = A.shape[1]
    dev = A.device
    G = torch.empty((R, N, N), dtype=torch.float64, device=dev)
    y = torch.empty((R, N), dtype=torch.float64, device=dev)
    sw = torch.empty((R,), dtype=torch.float64, device=dev)
    npts = torch.empty((R,), dtype=torch.int32, device=dev)
    Wm = torch.empty((R, P), dtype=torch.float64, device=dev) if want_masked else None
    bm = torch.empty((R, P), dtype=torch.float64, device=dev) if want_masked else None
    _native.check(_native.lib().vi_normal_eq_batched(
        A.data_ptr(), value.data_ptr(), error.data_ptr() if error is not None else None,
        weight.data_ptr() if weight is not None else None, R, P, N, mode,
        G.data_ptr(), y.data_ptr(), sw.data_ptr(), npts.data_ptr(),
        Wm.data_ptr() if want_masked else None, bm.data_ptr() if want_masked else None, _cuda_stream(dev)))
    return G, y, sw, npts, Wm, bm


_ws_cache = {}


def _workspace(dev, R, P, N, nreg, systems):
    """Scratch for vi_fit_batched, cached per (device, shape) so repeated fits do not re-allocate."""
    import torch
    need = C.c_int64(0)
    _native.check(_native.lib().vi_fit_workspace_bytes(R, P, N, nreg, systems, C.byref(need)))
    key = str(dev)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need.value:
        _ws_cache.pop(key, None)
        ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    return ws


def fit_batch_device(At, Wm, bm, G, y, npts, regs, method, systems=0, want_cov=False, A=None, cov_out=None,
                     want_trace=False):
    """vi_fit_batched on device tensors.  regs: (nreg,N,N) tensor or None.
    cov_out: optional (R,N,N) tensor receiving the covariance -- a CUDA tensor, or a PINNED host tensor (the
    library then streams the covariance out chunk by chunk on a side stream while it computes the next chunk)."""
    import torch
    R, N = y.shape
    P = At.shape[1]
    dev = y.device
    nreg = 0 if regs is None else regs.shape[0]
    meth = _native.METHOD_NONE if nreg == 0 else method
    Cf = torch.empty((R, N), dtype=torch.float64, device=dev)
    dC = cov_out
    if want_cov and dC is None:
        dC = torch.empty((R, N, N), dtype=torch.float64, device=dev)
    if dC is not None and not dC.is_cuda and not dC.is_pinned():
        raise ValueError("cov_out on the host must be pinned memory")
    chi2 = torch.empty((R,), dtype=torch.float64, device=dev)
    lam = torch.zeros((R, max(nreg, 1)), dtype=torch.float64, device=dev)
    rank = torch.zeros((R,), dtype=torch.int32, device=dev)
    status = torch.zeros((R,), dtype=torch.int32, device=dev)
    if R == 0:
        return Cf, dC, chi2, lam[:, :nreg], rank, status, 0
    ws = _workspace(dev, R, P, N, nreg, systems)
    nsolve = C.c_int64(0)
    _native.check(_native.lib().vi_fit_batched(
        At.data_ptr(), A.data_ptr() if A is not None else None, Wm.data_ptr(), bm.data_ptr(), G.data_ptr(),
        y.data_ptr(), npts.data_ptr(),
        R, P, N, regs.data_ptr() if nreg else None, nreg, meth,
        Cf.data_ptr(), dC.data_ptr() if dC is not None else None, chi2.data_ptr(), lam.data_ptr(),
        rank.data_ptr(), status.data_ptr(), C.byref(nsolve), ws.data_ptr(), ws.numel(), _cuda_stream(dev)))
    out = (Cf, dC, chi2, lam[:, :nreg], rank, status, nsolve.value)
    if want_trace and meth == _native.METHOD_CHI2:
        U = R * nreg
        table = torch.empty((U, _native.VI_NALPHA), dtype=torch.float64, device=dev)
        nu = torch.empty((U,), dtype=torch.float64, device=dev)
        klo = torch.empty((U,), dtype=torch.int32, device=dev)
        kdone = torch.empty((U,), dtype=torch.int32, device=dev)
        _native.check(_native.lib().vi_fit_search_trace(ws.data_ptr(), ws.numel(), R, P, nreg, table.data_ptr(),
                                                        nu.data_ptr(), klo.data_ptr(), kdone.data_ptr(),
                                                        _cuda_stream(dev)))
        out = out + ({"table": table, "nu": nu, "k_lo": klo, "kdone": kdone},)
    return out


def fit_records(model, lat, lon, alt, value, error, reg_matrices=None, method='chi2',
                ne_mode=_native.NE_FAST, weight=None, device=None, rec_batch=16384, systems=0,
                want_cov=False, to_host=True, want_trace=False, h2d_chunks=4):
    """Fit every record (reference interpolate.py:511-579).

    lat/lon/alt: (P,) NaN-altitude-filtered gate coordinates; value/error: (R,P) with NaN
    marking invalid gates; reg_matrices: list of (N,N) arrays in REGULARIZATION_LIST order.
    weight: optional (R,P) precomputed error**-2 (bit-exact parity with numpy's pow).

    to_host=True returns numpy arrays that are VIEWS of pinned host buffers owned by the result (no second
    host copy); the covariance is streamed out by the library while it is being computed.  Host inputs are
    uploaded in `h2d_chunks` slices per record batch on a side stream, each slice's normal equations
    starting as soon as it has landed."""
    import torch
    if method not in METHODS:
        raise ValueError(f"REGULARIZATION_METHOD {method!r} is not available on the device path (chi2, gcv)")
    dev = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    f64 = lambda a: a.to(dev, torch.float64) if isinstance(a, torch.Tensor) else \
        torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev, non_blocking=True)
    la, lo, al = f64(np.ravel(lat)), f64(np.ravel(lon)), f64(np.ravel(alt))
    P, N = la.numel(), model.nbasis
    A = torch.empty((P, N), dtype=torch.float64, device=dev)
    At = torch.empty((N, P), dtype=torch.float64, device=dev)
    model.basis_device(la, lo, al, out=A, out_t=At)
    regs = None
    if reg_matrices is not None and len(reg_matrices) > 0:
        regs = f64(np.stack([np.asarray(m, dtype=np.float64) for m in reg_matrices]))
    nreg = 0 if regs is None else regs.shape[0]
    R = value.shape[0]
    host = _HostResult(R, N, nreg, want_cov) if to_host else None
    outs, traces = [], []
    nsolve = 0
    main = torch.cuda.current_stream(dev)
    for r0 in range(0, R, rec_batch):
        r1 = min(R, r0 + rec_batch)
        G, y, npts, Wm, bm = _upload_and_normal_equations(A, value, error, weight, r0, r1, ne_mode, dev, h2d_chunks)
        cov_out = host.Covariance[r0:r1] if (host is not None and want_cov) else None
        o = fit_batch_device(At, Wm, bm, G, y, npts, regs, METHODS[method], systems, want_cov, A=A, cov_out=cov_out,
                             want_trace=want_trace)
        nsolve += o[6]
        if want_trace and len(o) > 7:
            traces.append(o[7])
        if host is not None:
            host.take(r0, r1, o)            # async D2H of the small per-record results
        else:
            outs.append(o[:6])
        del G, y, Wm, bm
    trace = None
    if traces:
        trace = {k: torch.cat([t[k] for t in traces]) for k in traces[0]}
        if to_host:
            trace = {k: v.cpu().numpy() for k, v in trace.items()}
    if host is not None:
        main.synchronize()
        return FitResult(*host.arrays(), nsolve=nsolve, trace=trace, device_small=host.device_small())
    if not outs:
        z = lambda *shape, dt=torch.float64: torch.empty(shape, dtype=dt, device=dev)
        return FitResult(z(0, N), z(0, N, N) if want_cov else None, z(0), z(0, nreg), z(0, dt=torch.int32),
                         z(0, dt=torch.int32), nsolve=0, trace=trace)
    cat = lambda i: (outs[0][i] if len(outs) == 1 else torch.cat([o[i] for o in outs])) \
        if outs[0][i] is not None else None
    return FitResult(*[cat(i) for i in range(6)], nsolve=nsolve, trace=trace)


def _upload_and_normal_equations(A, value, error, weight, r0, r1, ne_mode, dev, h2d_chunks):
    """value/error[r0:r1] -> device, normal equations (K2).  Host inputs go up in slices on a side stream; the
    normal equations of slice i run on the main stream while slice i+1 is in flight."""
    import torch
    P, N = A.shape
    R = r1 - r0
    G = torch.empty((R, N, N), dtype=torch.float64, device=dev)
    y = torch.empty((R, N), dtype=torch.float64, device=dev)
    npts = torch.empty((R,), dtype=torch.int32, device=dev)
    Wm = torch.empty((R, P), dtype=torch.float64, device=dev)
    bm = torch.empty((R, P), dtype=torch.float64, device=dev)
    if R == 0:
        return G, y, npts, Wm, bm
    on_dev = isinstance(value, torch.Tensor) and value.is_cuda
    nchunk = 1 if on_dev else max(1, min(h2d_chunks, R // 256))
    main = torch.cuda.current_stream(dev)
    side = _side_stream(dev) if nchunk > 1 else main
    bounds = [r0 + (R * i) // nchunk for i in range(nchunk + 1)]
    lib = _native.lib()

    def up(a, lo, hi):
        if a is None:
            return None
        s = a[lo:hi]
        if isinstance(s, torch.Tensor):
            return s.to(dev, torch.float64, non_blocking=True)
        return torch.from_numpy(np.ascontiguousarray(s, dtype=np.float64)).to(dev, non_blocking=True)

    staged = []
    for i in range(nchunk):
        lo, hi = bounds[i], bounds[i + 1]
        with torch.cuda.stream(side):
            v, e, w = up(value, lo, hi), up(error, lo, hi), up(weight, lo, hi)
            ev = torch.cuda.Event()
            ev.record(side)
        staged.append((lo, hi, v, e, w, ev))
    for lo, hi, v, e, w, ev in staged:
        main.wait_event(ev)
        a, b = lo - r0, hi - r0
        for t in (v, e, w):
            if t is not None:
                t.record_stream(main)
        _native.check(lib.vi_normal_eq_batched(
            A.data_ptr(), v.data_ptr(), e.data_ptr() if e is not None else None,
            w.data_ptr() if w is not None else None, hi - lo, P, N, ne_mode,
            G[a:b].data_ptr(), y[a:b].data_ptr(), None, npts[a:b].data_ptr(), Wm[a:b].data_ptr(), bm[a:b].data_ptr(),
            main.cuda_stream))
    return G, y, npts, Wm, bm


_side = {}


def _side_stream(dev):
    import torch
    key = str(dev)
    if key not in _side:
        _side[key] = torch.cuda.Stream(dev)
    return _side[key]


# ---- pinned result buffers ---------------------------------------------------------------------
# Results come back as numpy VIEWS of pinned host memory (no second, pageable copy).  Pinning is expensive
# (cudaHostAlloc of 1.7 GB ~ 0.5 s), so buffers are recycled: a buffer returns to the free list when the tensor the
# numpy views hang off is garbage-collected, i.e. when no result array refers to it any more.
_free_pinned = {}      # (dtype, numel) -> [free pinned 1-D tensors]
_NP_OF = {}


def _alloc_pinned(numel, dtype):
    import torch
    _NP_OF.update({torch.float64: np.float64, torch.int32: np.int32})
    return torch.empty((numel,), dtype=dtype).pin_memory()


class _Lease:
    """Keeps one pinned buffer checked out for as long as any numpy array created from it (or any view of such an
    array) is alive: numpy holds a reference to the object whose __array_interface__ it wrapped."""

    def __init__(self, flat, shape):
        self.flat = flat                       # the pinned storage (1-D tensor)
        self.tensor = flat[: int(np.prod(shape)) if len(shape) else 1].view(shape)
        self.__array_interface__ = {"version": 3, "shape": tuple(int(x) for x in shape), "data": (flat.data_ptr(), False),
                                    "typestr": np.dtype(_NP_OF[flat.dtype]).str}

    def numpy(self):
        return np.asarray(self)


def _pinned_buffer(shape, dtype):
    """-> _Lease over a pinned buffer of `shape`; the buffer returns to the free list when the lease dies."""
    import weakref
    numel = int(np.prod(shape)) if len(shape) else 1
    key = (dtype, numel)
    pool = _free_pinned.setdefault(key, [])
    flat = pool.pop() if pool else _alloc_pinned(max(numel, 1), dtype)
    lease = _Lease(flat, shape)
    weakref.finalize(lease, pool.append, flat)
    return lease


class _HostResult:
    """Pinned destination of one fit_records call."""

    def __init__(self, R, N, nreg, want_cov):
        import torch
        mk = lambda shape, dt: _pinned_buffer(shape, dt)
        self._lease = {"Coeffs": mk((R, N), torch.float64),
                       "Covariance": mk((R, N, N), torch.float64) if want_cov else None,
                       "chi_sq": mk((R,), torch.float64), "reg_params": mk((R, nreg), torch.float64),
                       "rank": mk((R,), torch.int32), "status": mk((R,), torch.int32)}
        self.Covariance = self._lease["Covariance"].tensor if want_cov else None
        self._dev = []

    def take(self, r0, r1, o):
        Cf, _, chi2, lam, rank, status = o[:6]       # the covariance was written by the library itself
        self._dev.append((Cf, chi2, lam, rank, status))
        L = self._lease
        L["Coeffs"].tensor[r0:r1].copy_(Cf, non_blocking=True)
        L["chi_sq"].tensor[r0:r1].copy_(chi2, non_blocking=True)
        if lam.shape[1]:
            L["reg_params"].tensor[r0:r1].copy_(lam, non_blocking=True)
        L["rank"].tensor[r0:r1].copy_(rank, non_blocking=True)
        L["status"].tensor[r0:r1].copy_(status, non_blocking=True)

    def device_small(self):
        import torch
        if not self._dev:
            return None
        keys = ("Coeffs", "chi_sq", "reg_params", "rank", "status")
        return {k: (self._dev[0][i] if len(self._dev) == 1 else torch.cat([d[i] for d in self._dev]))
                for i, k in enumerate(keys)}

    def arrays(self):
        return [self._lease[k].numpy() if self._lease[k] is not None else None
                for k in ("Coeffs", "Covariance", "chi_sq", "reg_params", "rank", "status")]
