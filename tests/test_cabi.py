"""The C-ABI library: loads, exports every symbol include/volinterp_b200.h declares, and the
product path fails loudly (no CPU fallback) when it is missing.  No compute calls: CPU only."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "volinterp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(vi_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from volumetricinterp_b200 import _native, build
    build.build()
    lib = C.CDLL(_native.LIB_PATH)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n
    # and the ctypes table binds exactly the declared compute entry points
    assert set(_native.SIGNATURES) | {"vi_version", "vi_last_error"} == set(names)
    lib.vi_version.restype = C.c_char_p
    assert b"sm_100a" in lib.vi_version()


def test_struct_layout_matches_header():
    from volumetricinterp_b200 import _native
    L = _native.VI_MAXL_MAX
    assert C.sizeof(_native.ShlParams) == 8 + 4 * 8 + L * 8 + 3 * L * L * 8


def test_no_cpu_fallback_when_library_missing(monkeypatch):
    from volumetricinterp_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libvolinterp_b200.so")
    with pytest.raises(_native.NativeLibraryMissing):
        _native.lib()


def test_error_reporting_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on the CPU box."""
    from volumetricinterp_b200 import _native
    lib = _native.lib()
    need = C.c_int64(0)
    assert lib.vi_fit_workspace_bytes(10, 100, 2000, 1, 0, C.byref(need)) == -4      # VI_EUNSUPPORTED
    assert b"1024" in lib.vi_last_error()
    assert lib.vi_fit_workspace_bytes(10, 100, 144, 1, 64, C.byref(need)) == 0
    assert need.value > 64 * 144 * 144 * 8


def test_product_never_imports_oracle():
    """The product package must not reach into oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "volumetricinterp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "ref_port" not in txt and "oracle" not in txt.replace("oracle/", "").lower() or f == "synth.py", f
