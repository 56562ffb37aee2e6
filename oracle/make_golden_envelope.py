"""TEST INFRASTRUCTURE -- reproducibility envelope of the reference at the rank-deficient orders, as a fixture.

For every record of the golden cases mid27, c1_144 and c3_500 (tests/golden/*.npz, produced by the UNMODIFIED reference):
the reference's algorithm (oracle/ref_port.py, bit-identical to the reference on these cases) is re-run
  * as shipped                                    ('gelsd': scipy.linalg.lstsq default driver, einsum normal equations)
  * with BLAS-order normal equations              ('blas':  a 1e-16 relative change of X)
  * with LAPACK gelss instead of gelsd            ('gelss': same rcond = eps, QR-iteration SVD instead of D&C)
and scale factor, bracket decade, lambda, rank, chi2 and the chi2(10^-k) table of each run are written to
tests/golden/envelope_<case>.json.  The GPU parity tests (tests/test_gpu_parity.py) require the CUDA path to land
inside what these three equally valid executions of the reference span.

    python oracle/make_golden_envelope.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity                      # noqa: E402
from conftest import load_golden   # noqa: E402


def main():
    for case in (sys.argv[1:] or ("mid27", "c1_144", "c3_500")):
        g = load_golden(case)
        out = {"case": case, "records": []}
        for r in range(g["value"].shape[0]):
            row = {}
            for order in ("einsum", "blas", "gelss"):
                o = parity.oracle_record(g["A"], g["value"][r], g["error"][r], g["regs"][0], g["reglist"][0], order)
                ok = np.isfinite(g["value"][r])
                row["gelsd" if order == "einsum" else order] = {
                    "status": o["status"], "sf": o["sf"], "k_lo": o["k_lo"], "lam": o["lam"], "rank": o["rank"],
                    "chi2": o["chi2"], "npts": o["npts"],
                    "table": [None if not np.isfinite(x) else float(x) for x in o["table"]],
                    "AC": None if o["AC"] is None else [float(x) for x in o["AC"]]}
            assert np.array_equal(parity.oracle_record(g["A"], g["value"][r], g["error"][r], g["regs"][0],
                                                       g["reglist"][0])["C"], g["Coeffs"][r], equal_nan=True)
            out["records"].append(row)
        with open(os.path.join(ROOT, "tests", "golden", f"envelope_{case}.json"), "w") as f:
            json.dump(out, f)
        print(case, [(x["gelsd"]["k_lo"], x["blas"]["k_lo"], x["gelss"]["k_lo"]) for x in out["records"]])


if __name__ == "__main__":
    main()
