import ctypes as C
import io
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    d = {k: g[k] for k in g.files}
    d["model_keys"] = json.loads(str(d["model_keys"]))
    d["default_keys"] = json.loads(str(d["default_keys"]))
    d["config_text"] = str(d["config_text"])
    d["reglist"] = [str(x) for x in d["reglist"]]
    d["regs"] = [d["reg_" + r] for r in d["reglist"]]
    return d


def oracle_model(g):
    """oracle (ref_port) model instance for a golden case."""
    import configparser
    import ref_port as rp
    cfg = configparser.ConfigParser()
    cfg.read_file(io.StringIO(g["config_text"]))
    return rp.model_from_config(cfg)


def product_model(g):
    import importlib
    name = g["model_keys"]["NAME"]
    m = importlib.import_module("volumetricinterp_b200.models." + name)
    return m.Model(io.StringIO(g["config_text"]))


@pytest.fixture(scope="session")
def harness():
    """TEST-ONLY CPU build of the device-agnostic headers (tests/cpu_harness.cpp)."""
    src = os.path.join(ROOT, "tests", "cpu_harness.cpp")
    so = os.path.join(ROOT, "tests", "_cpu_harness.so")
    deps = [src] + [os.path.join(ROOT, "volumetricinterp_b200", "csrc", f)
                    for f in ("vi_math.h", "vi_tql.h", "vi_brent.h", "vi_tridiag.h", "vi_tridiag_packed.h", "vi_ne_split.h",
                              "vi_nm.h", "vi_simt.h", "vi_band.h", "vi_chase.h")] + \
        [os.path.join(ROOT, "include", "volinterp_b200.h"), os.path.join(ROOT, "tests", "cuda_emu.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "tests"), "-o", so, src])
    return C.CDLL(so)


def dptr(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from volumetricinterp_b200 import _native
    _native.lib()     # fails loudly if the library is missing: no fallback
    return torch.device("cuda", 0)
