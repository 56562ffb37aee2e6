"""Device driver of the fit hot path: records in, coefficients out.

Python stays a thin host layer (device memory through torch, kernels through the
C ABI in _native.py).  One call fits all R records of a file:

    K1  design matrix A (P x N) and its transpose, once per file      (vi_basis_*)
    K2  masked weights + normal equations for a batch of records      (vi_normal_eq_batched)
    K3  chi^2 = nu search, final solve, chi^2                         (vi_fit_batched)

which replaces the reference's record loop, interpolate.py:511-574.
"""
import ctypes as C
import os
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


def fit_batch_device(At, Wm, bm, G, y, npts, regs, method, systems=0, want_cov=False, A=None):
    """vi_fit_batched on device tensors.  regs: (nreg,N,N) tensor or None."""
    import torch
    R, N = y.shape
    P = At.shape[1]
    dev = y.device
    nreg = 0 if regs is None else regs.shape[0]
    meth = _native.METHOD_NONE if nreg == 0 else method
    Cf = torch.empty((R, N), dtype=torch.float64, device=dev)
    dC = torch.empty((R, N, N), dtype=torch.float64, device=dev) if want_cov else None
    chi2 = torch.empty((R,), dtype=torch.float64, device=dev)
    lam = torch.zeros((R, max(nreg, 1)), dtype=torch.float64, device=dev)
    rank = torch.zeros((R,), dtype=torch.int32, device=dev)
    status = torch.zeros((R,), dtype=torch.int32, device=dev)
    ws = _workspace(dev, R, P, N, nreg, systems)
    nsolve = C.c_int64(0)
    _native.check(_native.lib().vi_fit_batched(
        At.data_ptr(), A.data_ptr() if A is not None else None, Wm.data_ptr(), bm.data_ptr(), G.data_ptr(),
        y.data_ptr(), npts.data_ptr(),
        R, P, N, regs.data_ptr() if nreg else None, nreg, meth,
        Cf.data_ptr(), dC.data_ptr() if want_cov else None, chi2.data_ptr(), lam.data_ptr(),
        rank.data_ptr(), status.data_ptr(), C.byref(nsolve), ws.data_ptr(), ws.numel(), _cuda_stream(dev)))
    return Cf, dC, chi2, lam[:, :nreg], rank, status, nsolve.value


def fit_records(model, lat, lon, alt, value, error, reg_matrices=None, method='chi2',
                ne_mode=_native.NE_FAST, weight=None, device=None, rec_batch=8192, systems=0,
                want_cov=False, to_host=True):
    """Fit every record (reference interpolate.py:511-579).

    lat/lon/alt: (P,) NaN-altitude-filtered gate coordinates; value/error: (R,P) with NaN
    marking invalid gates; reg_matrices: list of (N,N) arrays in REGULARIZATION_LIST order.
    weight: optional (R,P) precomputed error**-2 (bit-exact parity with numpy's pow)."""
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
    R = value.shape[0]
    outs = []
    nsolve = 0
    for r0 in range(0, R, rec_batch):
        r1 = min(R, r0 + rec_batch)
        v, e = f64(value[r0:r1]), f64(error[r0:r1])
        w = f64(weight[r0:r1]) if weight is not None else None
        G, y, _, npts, Wm, bm = normal_equations_device(A, v, e, w, ne_mode)
        Cf, dC, chi2, lam, rank, status, ns = fit_batch_device(At, Wm, bm, G, y, npts, regs, METHODS[method],
                                                                systems, want_cov, A=A)
        nsolve += ns
        outs.append((Cf, dC, chi2, lam, rank, status))
    cat = lambda i: (outs[0][i] if len(outs) == 1 else torch.cat([o[i] for o in outs])) \
        if outs and outs[0][i] is not None else None
    res = [cat(i) for i in range(6)]
    if to_host:
        res = _to_host(res, dev)
    return FitResult(*res, nsolve=nsolve)


_pinned = {}


def _to_host(tensors, dev):
    """Device results -> numpy through cached pinned staging buffers (one async copy each, one sync).
    The returned arrays own their memory (copied out of the staging buffers)."""
    import torch
    staged = []
    for i, t in enumerate(tensors):
        if t is None:
            staged.append(None)
            continue
        key = (i, t.dtype)
        buf = _pinned.get(key)
        if buf is None or buf.numel() < t.numel():
            buf = torch.empty((t.numel(),), dtype=t.dtype).pin_memory()
            _pinned[key] = buf
        view = buf[: t.numel()].view(t.shape)
        view.copy_(t, non_blocking=True)
        staged.append(view)
    torch.cuda.synchronize(dev)
    # pageable copies that own their memory; torch's CPU copy is multi-threaded (the covariance block is
    # 1.7 GB for 10k records: a single-threaded numpy copy with its page faults costs more than the PCIe transfer)
    out = []
    for v in staged:
        if v is None:
            out.append(None)
            continue
        o = torch.empty(v.shape, dtype=v.dtype)
        _parallel_copy(o, v)
        out.append(o.numpy())
    return out


_copy_pool = None


def _parallel_copy(dst, src, min_bytes=64 << 20):
    """dst.copy_(src) for host tensors, split over a few Python threads (torch's copy releases the GIL).  Does not
    depend on OMP_NUM_THREADS, which torchrun sets to 1 for every rank."""
    global _copy_pool
    import torch
    nbytes = src.numel() * src.element_size()
    if nbytes < min_bytes or src.dim() == 0 or src.shape[0] < 2:
        dst.copy_(src)
        return
    import concurrent.futures
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)
    nthr = max(1, min(8, (os.cpu_count() or 1) // max(1, local_world), src.shape[0]))
    if nthr == 1:
        dst.copy_(src)
        return
    if _copy_pool is None:
        _copy_pool = concurrent.futures.ThreadPoolExecutor(max_workers=8)
    step = (src.shape[0] + nthr - 1) // nthr
    futs = [_copy_pool.submit(lambda a=a: dst[a:a + step].copy_(src[a:a + step])) for a in range(0, src.shape[0], step)]
    for f in futs:
        f.result()
