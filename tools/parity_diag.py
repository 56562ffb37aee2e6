"""GPU box: tier-3 parity protocol (oracle/parity.py) on the rank-deficient golden cases, written to
gpurun_out/parity_<name>.json.  The oracle rows are the reference's algorithm (bit-identical to the unmodified
reference's coefficients on these fixtures, tests/test_oracle.py)."""
import io, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import parity
from conftest import load_golden, product_model
from volumetricinterp_b200 import fit

dev = torch.device("cuda", 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
for name in sys.argv[1:] or ["mid27", "c1_144"]:
    g = load_golden(name)
    model = product_model(g)
    res = fit.fit_records(model, g["lat"], g["lon"], g["alt"], g["value"], g["error"], g["regs"], "chi2", device=dev,
                          want_trace=True, want_cov=True)
    R = g["value"].shape[0]
    gpu = [parity.gpu_record(res, r, res.trace) for r in range(R)]
    ref = [parity.oracle_record(g["A"], g["value"][r], g["error"][r], g["regs"][0], g["reglist"][0]) for r in range(R)]
    env = [parity.oracle_record(g["A"], g["value"][r], g["error"][r], g["regs"][0], g["reglist"][0], "blas") for r in range(R)]
    alt = [parity.oracle_record(g["A"], g["value"][r], g["error"][r], g["regs"][0], g["reglist"][0], "gelss") for r in range(R)]
    out = parity.compare(gpu, ref, env, g["A"], g["value"], alt)
    out["golden_C_bit_identical_to_oracle"] = [bool(np.array_equal(ref[r]["C"], g["Coeffs"][r], equal_nan=True)) for r in range(R)]
    # decade by decade: GPU table vs the golden trace of the UNMODIFIED reference
    if "trace" in g:
        rows = []
        for r in range(R):
            tr = g["trace"][r]
            tr = tr[np.isfinite(tr[:, 0])]
            tab_ref = parity.rp._decade_table([tuple(x) for x in tr])
            tab_gpu = np.asarray(res.trace["table"][r])
            both = np.isfinite(tab_ref) & np.isfinite(tab_gpu)
            rel = np.abs(tab_gpu[both] - tab_ref[both]) / np.abs(tab_ref[both])
            n = int(np.isfinite(g["value"][r]).sum())
            rows.append({"record": r, "decades": int(both.sum()), "max_rel": float(rel.max()) if rel.size else None,
                         "worst_decade": int(np.flatnonzero(both)[rel.argmax()]) if rel.size else None,
                         "chi2_over_n_ref": (tab_ref[both] / n).round(6).tolist(),
                         "chi2_over_n_gpu": (tab_gpu[both] / n).round(6).tolist(),
                         "k_lo_gpu": int(res.trace["k_lo"][r]), "nu_gpu": float(res.trace["nu"][r]), "npts": n})
        out["table_vs_reference_trace"] = rows
    if res.Covariance is not None and "Covariance_diag" in g:
        out["cov_diag_rel"] = [float(np.nanmax(np.abs(np.diag(res.Covariance[r]) - g["Covariance_diag"][r])) /
                                     np.nanmax(np.abs(g["Covariance_diag"][r]))) if np.isfinite(g["Covariance_diag"][r]).all() else None
                               for r in range(R)]
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"parity_{name}.json"), "w"), indent=1, default=float)
    for k in ("gpu_vs_reference", "reference_vs_itself_blas_order", "reference_vs_itself_gelss_driver",
              "gpu_vs_reference_with_gelss_driver"):
        print(name, k, json.dumps(out[k]))
    for p in out["per_record"]:
        print("  ", json.dumps({k: p[k] for k in ("record", "ref", "gpu", "env_blas", "env_gelss")}, default=float))
