"""Build libvolinterp_b200.so in-tree with nvcc for sm_100a.

    python -m volumetricinterp_b200.build [--force]

One shared library, C ABI (include/volinterp_b200.h).  basis.cu is compiled with
-fmad=false so the per-point special-function arithmetic runs in the order
written in csrc/vi_math.h (the order scipy executes it in on the CPU).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(ROOT, "include")
# VI_LIB_VARIANT=<tag> (with VI_EXTRA_NVCC=-D...) builds an A/B variant next to the product library
_TAG = os.environ.get("VI_LIB_VARIANT", "")
OUT = os.path.join(HERE, "libvolinterp_b200%s.so" % (("_" + _TAG) if _TAG else ""))
OBJ = os.path.join(HERE, "build" + (("_" + _TAG) if _TAG else ""))

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INC, "-I", CSRC, "--expt-relaxed-constexpr"]
COMMON += os.environ.get("VI_EXTRA_NVCC", "").split()      # e.g. -DVI_TRP_PROFILE (phase timers of k_tridiag_packed)
UNITS = {          # file -> extra flags
    "abi.cu": [],
    "basis.cu": ["-fmad=false"],
    "normal_eq.cu": [],
    "estimate_gemm.cu": [],
    "fit.cu": [],
    "probe.cu": [],
}


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(INC, "volinterp_b200.h"))
    objs = []
    procs = []
    for src, extra in UNITS.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [_nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
            procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            print(" ".join(cmd))
            print(out)
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(OUT, objs):
        tmp = OUT + ".tmp"                 # link aside, then rename: the library is never seen half-written
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", tmp] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
        os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
