"""TEST INFRASTRUCTURE — runs the UNMODIFIED reference numerical code from
/root/reference (never copied into this repo) to produce golden vectors.

The reference is pure Python (numpy/scipy) but cannot be imported as shipped in
this image: `tables`, `h5py`, `pymap3d`, `matplotlib`, `cartopy` are absent.
`oracle/shims/` provides empty stubs for the I/O/plot imports and a one-function
`pymap3d.geodetic2ecef` (WGS84 closed form).  With those on sys.path every
function of SURVEY.md §8 rows A1-A15 executes unmodified; only the HDF5
reader/writer (`read_datafile`, `saveh5`, `loadh5`) is bypassed by assigning
arrays on the instance.

Must run in a separate interpreter (the reference hard-codes
`package='volumetricinterp'`, interpolate.py:61 / estimate.py:49):

    python oracle/run_reference.py   # only through oracle/make_golden.py

/root/reference exists only in the build container; nothing under tests/ -m gpu,
smoke() or bench.py may call this at run time.
"""
import io
import os
import sys
import tempfile
import datetime as dt

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("VI_REFERENCE_ROOT", "/root/reference")


def _activate():
    for p in (os.path.join(HERE, "shims"), REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def write_config(path, model, default=None):
    """Write an INI with the reference's keys (example_config.ini)."""
    d = dict(PARAM="dens", FILENAME="synthetic.h5", OUTPUTFILENAME="out.h5",
             REGULARIZATION_LIST="curvature", REGULARIZATION_METHOD="chi2",
             ERRLIM="1e10,1e13", GOODFITCODE="1,2,3,4", CHI2LIM="0.1,10")
    d.update(default or {})
    m = dict(NAME="sphharmlag", MAXK=4, MAXL=6, CAP_LIM=10, MAX_Z_INT="INF",
             LATCP=78, LONCP=262, EPS=100000.0, LATRANGE="74,80",
             LONRANGE="260,285", ALTRANGE="100,600", NUMGRIDPNT=7)
    m.update(model or {})
    with open(path, "w") as f:
        f.write("[DEFAULT]\n")
        for k, v in d.items():
            f.write(f"{k} = {v}\n")
        f.write("\n[MODEL]\n")
        for k, v in m.items():
            f.write(f"{k} = {v}\n")
    return path


def config_text(model, default=None):
    with tempfile.NamedTemporaryFile("w+", suffix=".ini", delete=False) as f:
        name = f.name
    write_config(name, model, default)
    txt = open(name).read()
    os.unlink(name)
    return txt


def reference_model(model, default=None):
    """Instantiate the reference's models.<NAME>.Model from config keys."""
    _activate()
    import importlib
    txt = config_text(model, default)
    name = (model or {}).get("NAME", "sphharmlag")
    m = importlib.import_module("volumetricinterp.models." + name)
    return m.Model(io.StringIO(txt)), txt


def reference_fit(model, default, utime, lat, lon, alt, value, error,
                  reg_matrices=None, trace=True):
    """Run Interpolate.calc_coeffs (interpolate.py:472-579) on in-memory arrays.

    lat/lon/alt: (P,) already NaN-altitude-filtered; value/error (R,P).
    Returns dict with Coeffs, Covariance, chi_sq, hull_vert, reg matrices and,
    if trace, the per-record (alpha, chi2-nu) evaluation trace and lambda."""
    _activate()
    import numpy as np
    from volumetricinterp.interpolate import Interpolate

    with tempfile.NamedTemporaryFile("w", suffix=".ini", delete=False) as f:
        cfg = f.name
    write_config(cfg, model, default)
    it = Interpolate(cfg)
    os.unlink(cfg)
    it.read_datafile = lambda fn: (utime.copy(), lat.copy(), lon.copy(),
                                   alt.copy(), value.copy(), error.copy())
    regs = {}
    if reg_matrices is not None:
        it.model.eval_reg_matricies = {k: (lambda v=v: v.copy()) for k, v in reg_matrices.items()}
    # record the regularisation matrices the run actually used
    orig = dict(it.model.eval_reg_matricies)
    def _wrap(name, fn):
        def g():
            regs[name] = fn()
            return regs[name]
        return g
    it.model.eval_reg_matricies = {k: _wrap(k, v) for k, v in orig.items()}

    traces, lams = [], []
    if trace:
        f0 = it.chi2objfunct
        def traced(alpha, A, b, W, reg_matrices, nu, reg):
            val = f0(alpha, A, b, W, reg_matrices, nu, reg)
            traces[-1].append((float(alpha), float(val), float(nu)))
            return val
        it.chi2objfunct = traced
        g0 = it.find_reg_param
        def traced_find(A, b, W, reg_matrices, method=None):
            traces.append([])
            out = g0(A, b, W, reg_matrices, method=method)
            lams.append([float(out[k]) for k in it.regularization_list])
            return out
        it.find_reg_param = traced_find

    import contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        it.calc_coeffs()
    out = dict(Coeffs=it.Coeffs, Covariance=it.Covariance, chi_sq=it.chi_sq,
               hull_vert=it.hull_vert, time=it.time,
               reglist=list(it.regularization_list))
    for k, v in regs.items():
        out["reg_" + k] = v
    if trace:
        out["lam"] = np.array(lams, dtype=float).reshape(len(lams), -1)
        ntr = max((len(t) for t in traces), default=0)
        tr = np.full((len(traces), ntr, 3), np.nan)
        for i, t in enumerate(traces):
            if t:
                tr[i, :len(t)] = np.array(t)
        out["trace"] = tr
        out["n_eval"] = np.array([len(t) for t in traces])
    return out


def reference_estimate(model, default, utime, Coeffs, hull_vert, when, lat, lon, alt,
                       check_hull=True, timetol=60.0, timeinterp=False):
    """Run Estimate.__call__ (estimate.py:75-123) with loadh5 bypassed."""
    _activate()
    import importlib
    import numpy as np
    from volumetricinterp.estimate import Estimate
    est = Estimate.__new__(Estimate)
    est.timetol, est.timeinterp = timetol, timeinterp
    est.Coeffs = Coeffs
    N = Coeffs.shape[1]
    est.Covariance = np.zeros((Coeffs.shape[0], 1, 1))
    est.time = utime
    est.hull_vert = hull_vert
    txt = config_text(model, default)
    name = (model or {}).get("NAME", "sphharmlag")
    m = importlib.import_module("volumetricinterp.models." + name)
    est.model = m.Model(io.StringIO(txt))
    return est(when, lat, lon, alt, check_hull=check_hull)


def unix2datetime(t):
    return dt.datetime.utcfromtimestamp(0) + dt.timedelta(seconds=float(t))
