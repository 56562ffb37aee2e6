"""TEST INFRASTRUCTURE — golden vectors for sphharmlag.Model.grad_basis (models/sphharmlag.py:148-184), produced by
executing the UNMODIFIED reference through oracle/run_reference.py.  Build container only.

    python oracle/make_golden_grad.py        # writes tests/golden/grad_basis.npz

grad_basis is not called anywhere on the reference's fit / Estimate path; the reference
function itself is the oracle.  Shape as the reference returns it: np.array(Ag).T = (npoints, 3, nbasis), axis 1 =
(z-hat, theta-hat, phi-hat).
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

CASES = {"g12": dict(NAME="sphharmlag", MAXK=3, MAXL=2, CAP_LIM=10),
         "g144": dict(NAME="sphharmlag", MAXK=4, MAXL=6, CAP_LIM=10)}

if __name__ == "__main__":
    out = {}
    lat2, lon2, alt2 = synth.make_geometry(9, 24, seed=4242)
    lat, lon, alt, _ = synth.flatten_valid(lat2, lon2, alt2)
    out["lat"], out["lon"], out["alt"] = lat, lon, alt
    for name, mk in CASES.items():
        model, cfg = rr.reference_model(mk, {})
        with np.errstate(all="ignore"):
            g = np.ascontiguousarray(model.grad_basis(lat, lon, alt))
            A = np.ascontiguousarray(model.basis(lat, lon, alt))
        out[name + "_grad"] = g
        out[name + "_A"] = A
        out[name + "_config_text"] = cfg
        out[name + "_model_keys"] = json.dumps(mk)
        print(name, g.shape, np.isfinite(g).all(), np.abs(g).max())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "grad_basis.npz"), **out)
