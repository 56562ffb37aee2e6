"""ctypes binding of the C-ABI library (include/volinterp_b200.h).

The library is hand-written CUDA for sm_100a (volumetricinterp_b200/csrc).  There
is NO CPU fallback: `lib()` raises if the shared object is missing or no CUDA
device is usable, and every product entry point goes through it.
"""
import ctypes as C
import os

import numpy as np

VI_MAXL_MAX = 16
VI_MAXK_MAX = 16
VI_NALPHA = 102

# record status codes (csrc/vi_brent.h)
ST_OK, ST_TOO_SMOOTH, ST_NO_ROOT, ST_NONFINITE, ST_NOCONV, ST_EMPTY = range(6)

_HERE = os.path.dirname(os.path.abspath(__file__))
# VI_LIB_VARIANT selects an A/B build made by `VI_LIB_VARIANT=<tag> python -m volumetricinterp_b200.build` (tuning only)
LIB_PATH = os.path.join(_HERE, "libvolinterp_b200%s.so" % (("_" + os.environ["VI_LIB_VARIANT"])
                                                          if os.environ.get("VI_LIB_VARIANT") else ""))


class ShlParams(C.Structure):
    """mirror of `struct vi_shl_params` (csrc/vi_math.h)."""
    _fields_ = [
        ("maxk", C.c_int32), ("maxl", C.c_int32),
        ("ct0", C.c_double), ("st0", C.c_double),
        ("kx", C.c_double), ("ky", C.c_double),
        ("nu", C.c_double * VI_MAXL_MAX),
        ("kvm", (C.c_double * VI_MAXL_MAX) * VI_MAXL_MAX),
        ("g1", (C.c_double * VI_MAXL_MAX) * VI_MAXL_MAX),
        ("g2", (C.c_double * VI_MAXL_MAX) * VI_MAXL_MAX),
    ]


class NativeLibraryMissing(RuntimeError):
    pass


_lib = None

_i32, _i64, _dbl, _ptr = C.c_int32, C.c_int64, C.c_double, C.c_void_p

# name -> argtypes; every function returns int status (0 ok).  Keep in sync with
# include/volinterp_b200.h (tests/test_cabi.py checks every header symbol).
_shl = C.POINTER(ShlParams)
SIGNATURES = {
    "vi_basis_sphharmlag": [_ptr, _ptr, _ptr, _i64, _shl, _ptr, _ptr, _ptr],
    "vi_grad_basis_sphharmlag": [_ptr, _ptr, _ptr, _i64, _shl, _ptr, _ptr],
    "vi_basis_radbasfun": [_ptr, _ptr, _ptr, _i64, _ptr, _i32, _dbl, _ptr, _ptr, _ptr],
    "vi_normal_eq_batched": [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "vi_fit_workspace_bytes": [_i32, _i32, _i32, _i32, _i64, C.POINTER(_i64)],
    "vi_solve_batched": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _dbl, _ptr, _ptr, _ptr, _ptr, _i64, _ptr],
    "vi_solve_cov_batched": [_ptr, _ptr, _ptr, _ptr, _ptr, _i64, _i32, _i32, _dbl, _ptr, _ptr, _ptr, _ptr, _ptr, _i64, _ptr],
    "vi_fit_batched": [_ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _ptr, _i32, _i32,
                       _ptr, _ptr, _ptr, _ptr, _ptr, _ptr, C.POINTER(_i64), _ptr, _i64, _ptr],
    "vi_fit_search_trace": [_ptr, _i64, _i32, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr],
    "vi_estimate_sphharmlag": [_ptr, _ptr, _ptr, _i64, _shl, _ptr, _i32, _ptr, _i32, _ptr, _ptr],
    "vi_estimate_radbasfun": [_ptr, _ptr, _ptr, _i64, _ptr, _i32, _dbl, _ptr, _i32, _ptr, _i32, _ptr, _ptr],
    "vi_estimate_workspace_bytes": [_i64, _i32, _i32, C.POINTER(_i64)],
    "vi_estimate_sphharmlag_many": [_ptr, _ptr, _ptr, _i64, _shl, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _i64, _ptr],
    "vi_estimate_radbasfun_many": [_ptr, _ptr, _ptr, _i64, _ptr, _i32, _dbl, _ptr, _i32, _ptr, _i32, _ptr, _ptr, _i64, _ptr],
    "vi_fit_host": [_ptr, _ptr, _ptr, _ptr, _i32, _i32, _i32, _ptr, _i32, _i32, _i32, _ptr, _ptr, _ptr, _ptr, _ptr, _ptr],
    "vi_estimate_sphharmlag_host": [_ptr, _ptr, _ptr, _i64, _shl, _ptr, _i32, _ptr, _i32, _ptr],
    "vi_fp64_peak_probe": [_i32, _i32, C.POINTER(_dbl), _ptr],
    "vi_profile_enable": [_i32],
    "vi_profile_reset": [],
    "vi_profile_read": [_ptr, _ptr, _i32],
    "vi_profile_counters": [C.POINTER(_i64), C.POINTER(_i64)],
}
PROFILE_KINDS = ("basis", "normal_eq", "tridiag", "tql", "apply", "chi2", "covariance", "estimate", "misc", "chase", "est_gemm")


def profile_read():
    """-> ({kind: device ms}, {kind: launches}) since the last vi_profile_reset."""
    n = len(PROFILE_KINDS)
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    check(lib().vi_profile_read(ms, cnt, n))
    return dict(zip(PROFILE_KINDS, list(ms))), dict(zip(PROFILE_KINDS, list(cnt)))
NE_STRICT, NE_FAST = 0, 1
METHOD_NONE, METHOD_CHI2, METHOD_GCV = 0, 1, 2


def lib():
    """Load libvolinterp_b200.so (built by `python -m volumetricinterp_b200.build`
    or __graft_entry__.build()).  Raises NativeLibraryMissing — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryMissing(
            f"{LIB_PATH} not found: build it with `python -m volumetricinterp_b200.build` "
            "(nvcc, sm_100a).  volumetricinterp_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    L.vi_last_error.restype = C.c_char_p
    L.vi_last_error.argtypes = []
    L.vi_version.restype = C.c_char_p
    L.vi_version.argtypes = []
    _lib = L
    return L


class NativeError(RuntimeError):
    pass


def check(status):
    if status != 0:
        raise NativeError(f"volinterp_b200 error {status}: {lib().vi_last_error().decode()}")


def fill_shl_params(maxk, maxl, ct0, st0, kx, ky, nu, kvm, g1, g2):
    if maxl > VI_MAXL_MAX or maxk > VI_MAXK_MAX:
        raise ValueError(f"MAXL<= {VI_MAXL_MAX} and MAXK <= {VI_MAXK_MAX} supported")
    p = ShlParams()
    p.maxk, p.maxl = int(maxk), int(maxl)
    p.ct0, p.st0, p.kx, p.ky = float(ct0), float(st0), float(kx), float(ky)
    for l in range(maxl):
        p.nu[l] = float(nu[l])
        for m in range(l + 1):
            p.kvm[l][m] = float(kvm[l][m])
            p.g1[l][m] = float(g1[l][m])
            p.g2[l][m] = float(g2[l][m])
    return p


_est_ws = {}


def estimate_workspace(device, npts, N, Rsel, slot=0):
    """Device scratch of vi_estimate_*_many (cached per device and slot, grown on demand; callers that keep two
    tiles in flight on two streams use two slots)."""
    import torch
    need = C.c_int64(0)
    check(lib().vi_estimate_workspace_bytes(int(npts), int(N), int(Rsel), C.byref(need)))
    key = (str(device), int(slot))
    ws = _est_ws.get(key)
    if ws is None or ws.numel() < need.value:
        _est_ws.pop(key, None)
        ws = torch.empty((need.value,), dtype=torch.uint8, device=device)
        _est_ws[key] = ws
    return ws
