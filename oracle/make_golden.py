"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by executing the UNMODIFIED
reference (/root/reference, through oracle/run_reference.py and the import shims)
on seeded synthetic AMISR-shaped inputs.  Build container only: the GPU box has
no /root/reference, the fixtures travel instead.

    python oracle/make_golden.py [case ...]

Each fixture holds the inputs (lat, lon, alt, value, error, utime, config keys),
the reference's design matrix, regularisation matrices and fit outputs (Coeffs,
Covariance, chi_sq, lambda, the (alpha, chi2-nu, nu) evaluation trace) and
Estimate outputs with the in-hull mask.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import run_reference as rr                     # noqa: E402
from volumetricinterp_b200 import synth       # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# name -> (model keys, default keys, nbeams, ngates, nrecords, keep_cov)
CASES = {
    # strict-parity tier: low order, full rank (SURVEY.md §8-c protocol 2)
    "lo8": (dict(NAME="sphharmlag", MAXK=2, MAXL=2, CAP_LIM=10), {}, 7, 30, 6, True),
    "lo12": (dict(NAME="sphharmlag", MAXK=3, MAXL=2, CAP_LIM=10), {}, 7, 30, 6, True),
    # two regularisers searched independently, then applied together
    "lo12_two": (dict(NAME="sphharmlag", MAXK=3, MAXL=2, CAP_LIM=10),
                 dict(REGULARIZATION_LIST="curvature,0thorder"), 7, 30, 4, True),
    # mid order: rank deficiency starts
    "mid27": (dict(NAME="sphharmlag", MAXK=3, MAXL=3, CAP_LIM=10), {}, 9, 40, 4, True),
    # C1 shape at the default order (example_config.ini): N = 144, 11 x 70
    "c1_144": (dict(NAME="sphharmlag", MAXK=4, MAXL=6, CAP_LIM=10), {}, 11, 70, 3, False),
    # gcv method (interpolate.py:263-351): leave-one-gate-out objective, Nelder-Mead; small on purpose
    # (the reference solves len(b) systems per objective evaluation)
    "lo8_gcv": (dict(NAME="sphharmlag", MAXK=2, MAXL=2, CAP_LIM=10), dict(REGULARIZATION_METHOD="gcv"), 5, 16, 3, True),
    # C3 order (BASELINE configs[2]): N = 500 (MAXK 5 x MAXL 10; CAP_LIM 11 keeps Gamma(nu + m + 1) finite),
    # the reference's real curvature matrix (9 minutes of QUADPACK), few gates and records to keep the run offline
    "c3_500": (dict(NAME="sphharmlag", MAXK=5, MAXL=10, CAP_LIM=11), {}, 9, 40, 2, False),
    # radbasfun: no regulariser exists (radbasfun.py:62) -> plain lstsq per record
    "rbf27": (dict(NAME="radbasfun", NUMGRIDPNT=3, EPS=300000.0, LATRANGE="74,80", LONRANGE="255,280",
                   ALTRANGE="100,600"), dict(REGULARIZATION_LIST=""), 9, 40, 4, True),
}


def build_case(name):
    model, default, nbeams, ngates, nrec, keep_cov = CASES[name]
    seed = sum(ord(c) for c in name)
    lat2, lon2, alt2 = synth.make_geometry(nbeams, ngates, seed=seed)
    lat, lon, alt, keep = synth.flatten_valid(lat2, lon2, alt2)
    ref_model, cfg_text = rr.reference_model(model, default)
    A = np.ascontiguousarray(ref_model.basis(lat, lon, alt))
    maxl = model.get("MAXL") if model["NAME"] == "sphharmlag" else None
    if model["NAME"] == "radbasfun":
        c_true = np.zeros(A.shape[1]); c_true[::5] = 2e11
        value, error, _ = synth.make_records(A, nrec, seed=seed + 1, c_true=c_true, noise_scale=0.85)
    else:
        value, error, _ = synth.make_records(A, nrec, seed=seed + 1, maxl=maxl, noise_scale=0.85)
    # one record with no chi^2 root: pure noise far above the errors -> every scale factor fails -> NaN record
    if nrec >= 4 and model["NAME"] == "sphharmlag" and default.get("REGULARIZATION_METHOD", "chi2") == "chi2":
        rng = np.random.default_rng(seed + 2)
        ok = np.isfinite(value[1])
        value[1, ok] = 5e11 * rng.standard_normal(ok.sum())
        error[1, ok] = 1.2e10
    utime = synth.make_unixtime(nrec)
    out = rr.reference_fit(model, default, utime, lat, lon, alt, value, error)
    fx = dict(lat=lat, lon=lon, alt=alt, value=value, error=error, utime=utime, A=A,
              Coeffs=out["Coeffs"], chi_sq=out["chi_sq"], hull_vert=out["hull_vert"],
              config_text=np.array(cfg_text), model_keys=np.array(json.dumps(model)),
              default_keys=np.array(json.dumps(default)), reglist=np.array(out["reglist"]))
    if keep_cov:
        fx["Covariance"] = out["Covariance"]
    else:
        fx["Covariance_diag"] = np.array([np.diag(c) for c in out["Covariance"]])
    for k in out:
        if k.startswith("reg_"):
            fx[k] = out[k]
    if "lam" in out:
        fx["lam"], fx["trace"], fx["n_eval"] = out["lam"], out["trace"], out["n_eval"]
    # Estimate: a small grid straddling the hull so the mask has both values
    rng = np.random.default_rng(seed + 3)
    nq = 60
    qlat = rng.uniform(np.nanmin(lat) + 0.5, np.nanmax(lat) - 0.5, nq)
    qlon = rng.uniform(np.nanmin(lon) + 2.0, np.nanmax(lon) - 2.0, nq)
    qalt = rng.uniform(150e3, 600e3, nq)
    good = [i for i in range(nrec) if np.all(np.isfinite(out["Coeffs"][i]))]
    rec = good[0]
    when = rr.unix2datetime(utime[rec].mean() + 5.0)
    est = rr.reference_estimate(model, default, utime, out["Coeffs"], out["hull_vert"], when,
                                qlat.reshape(6, 10), qlon.reshape(6, 10), qalt.reshape(6, 10))
    fx.update(q_lat=qlat.reshape(6, 10), q_lon=qlon.reshape(6, 10), q_alt=qalt.reshape(6, 10),
              q_record=np.array(rec), q_time=np.array(utime[rec].mean() + 5.0), q_out=est)
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **fx)
    print(name, "P=%d N=%d R=%d" % (A.shape[0], A.shape[1], nrec), "lam=", fx.get("lam", np.zeros(0)).ravel(),
          "nan records:", [i for i in range(nrec) if i not in good], "inside:", int(np.isfinite(est).sum()), "/", nq,
          flush=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(CASES)):
        build_case(n)
