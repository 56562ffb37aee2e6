"""profiles/r02_parity.md from the parity artefacts of a GPU run:
    gpurun_out/parity_<case>.json   (tools/parity_diag.py: golden cases of the unmodified reference)
    a bench.py JSON line            (its `parity` block: first records of the benchmarked workload)
usage: python tools/make_parity_md.py <bench.json> > profiles/r02_parity.md"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def fmt(x, p=2):
    if x is None:
        return "-"
    if isinstance(x, float):
        return f"{x:.{p}e}" if (abs(x) < 1e-2 or abs(x) >= 1e4) and x != 0 else f"{x:.4g}"
    return str(x)


def summary_table(block, keys):
    rows = [("quantity",) + tuple(k[1] for k in keys)]
    def cell(s, name):
        v = s.get(name)
        if v is None:
            return "-"
        if "agree" in v:
            return f"{v['agree']} / {v['of']}"
        return f"median {fmt(v['median'])}, max {fmt(v['max'])} (n={v['n']})"
    names = [("status_identical", "record status identical (NaN / lambda = 0 / root)"),
             ("scale_factor_identical", "scale factor identical"),
             ("bracket_decade_identical", "bracket decade identical"),
             ("abs_dlog10_lambda", "|log10 lambda - log10 lambda_ref|"),
             ("rank_diff_abs", "|rank - rank_ref|"),
             ("AC_rel_diff", "max|A.C - A.C_ref| / max|A.C_ref|"),
             ("chi2_rel_diff", "|chi2 - chi2_ref| / chi2_ref"),
             ("table_rel_diff", "chi2(10^-k) table, max rel. diff over the decades both evaluated"),
             ("C_rel_diff", "max|C - C_ref| / max|C_ref|")]
    out = ["| " + " | ".join(rows[0]) + " |", "|" + "---|" * len(rows[0])]
    for key, label in names:
        out.append("| " + label + " | " + " | ".join(cell(block[k[0]], key) if k[0] in block else "-" for k in keys) + " |")
    return "\n".join(out)


KEYS = [("gpu_vs_reference", "GPU vs reference (gelsd, as shipped)"),
        ("gpu_vs_reference_with_gelss_driver", "GPU vs reference with LAPACK gelss"),
        ("reference_vs_itself_blas_order", "envelope 1: reference vs itself, BLAS-order A^T W A"),
        ("reference_vs_itself_gelss_driver", "envelope 2: reference vs itself, gelss instead of gelsd")]


def main():
    print("# Round 2 — tier-3 parity at the rank-deficient orders (SURVEY.md §8-c item 3)\n")
    print("Protocol (`oracle/parity.py`): per record, the CUDA fit against the oracle (`oracle/ref_port.py`, bit-identical to the\n"
          "unmodified reference's coefficients, lambda and evaluation trace on every golden case — `tests/test_oracle.py`), next to\n"
          "two executions of the *reference against itself*: with its normal equations summed in BLAS order instead of\n"
          "`np.einsum` order (a 1e-16 relative change of X), and with `scipy.linalg.lstsq` running LAPACK `gelss` instead of\n"
          "`gelsd` (same `rcond = eps`).  All three are equally valid executions of `interpolate.py:152-218, 456-462`; how far\n"
          "they move is how far the reference's own answer is defined.\n")
    print("Reading: wherever the three reference executions agree (record status, scale factor, bracket decade) the GPU\n"
          "agrees with them on every record of the golden cases and all but a few per cent of the benchmark records (listed\n"
          "below); where they disagree (lambda, rank, fitted densities, the chi2 table past the decade where the eps-truncation\n"
          "sets in) the GPU lies inside their spread — and is closest to the `gelss` execution, i.e. the symmetric\n"
          "eigen-solver of the CUDA path takes the same rank decisions as LAPACK's QR-iteration SVD.\n")
    for case in ("mid27", "c1_144"):
        p = os.path.join(ROOT, "gpurun_out", f"parity_{case}.json")
        if not os.path.exists(p):
            continue
        d = json.load(open(p))
        n = {"mid27": "N = 27 (MAXK 3, MAXL 3), 9 x 40 gates, 4 records", "c1_144": "N = 144 (example_config.ini order), 11 x 70 gates, 3 records"}[case]
        print(f"## Golden case `{case}` — {n}; goldens produced by the UNMODIFIED reference\n")
        print(summary_table(d, KEYS) + "\n")
        print("| record | reference (sf, decade, lambda, rank) | GPU | reference, BLAS order | reference, gelss |")
        print("|---|---|---|---|---|")
        for r in d["per_record"]:
            c = lambda x: "-" if x is None else f"{fmt(x.get('sf'))}, {x.get('k_lo')}, {fmt(x.get('lam'), 6)}, {x.get('rank')}"
            print(f"| {r['record']} | {c(r['ref'])} | {c(r['gpu'])} | {c(r.get('env_blas'))} | {c(r.get('env_gelss'))} |")
        if "table_vs_reference_trace" in d:
            print("\nχ²(10^-k) table of the GPU against the golden `(alpha, chi2 - nu)` trace of the unmodified reference, decade by decade:\n")
            print("| record | decades compared | max rel. diff, k <= 20 | max rel. diff, all decades | at decade | GPU bracket decade |")
            print("|---|---|---|---|---|---|")
            import numpy as np
            for t in d["table_vs_reference_trace"]:
                a, b = np.array(t["chi2_over_n_ref"]), np.array(t["chi2_over_n_gpu"])
                rel = np.abs(a - b) / np.abs(a)
                print(f"| {t['record']} | {t['decades']} | {fmt(float(rel[:21].max()))} | {fmt(t['max_rel'])} | {t['worst_decade']} | {t['k_lo_gpu']} |")
        if d.get("cov_diag_rel"):
            print("\ncovariance diagonal, max|dC_ii − dC_ii,ref| / max|dC_ii,ref| per record: " + ", ".join(fmt(x) for x in d["cov_diag_rel"]) +
                  " (at N = 144 the reference's own covariance is rounding noise: its golden diagonal has negative entries)")
        print()
    if len(sys.argv) > 1 and os.path.exists(sys.argv[1]):
        line = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
        par = line.get("parity")
        if par:
            cfgw = line["config"]
            print(f"## Benchmark workload — first {par['records_compared']} records of the timed batch "
                  f"({cfgw['gates']} gates, N = {cfgw['nbasis']}, noise scale {cfgw.get('noise_scale')}, {cfgw.get('signal_terms')} signal terms)\n")
            print(f"The rows the parity run fitted alone are bit-identical to the same rows of the full {cfgw['records_per_gpu']}-record batch: "
                  f"**{par['gpu_rows_identical_to_the_full_batch']}**.\n")
            print(summary_table(par, KEYS) + "\n")
            bad = [r for r in par["per_record"] if not (r["gpu_vs_ref"].get("status_same", True) and r["gpu_vs_ref"].get("k_lo_same", True)
                                                          and r["gpu_vs_ref"].get("sf_same", True))]
            if bad:
                print("Records on which the GPU's bracket differs from the reference's:\n")
                print("| record | reference (sf, decade, lambda, rank) | GPU | reference, BLAS order | reference, gelss |")
                print("|---|---|---|---|---|")
                for r in bad:
                    c = lambda x: "-" if x is None else f"{fmt(x.get('sf'))}, {x.get('k_lo')}, {fmt(x.get('lam'), 6)}, {x.get('rank')}"
                    print(f"| {r['record']} | {c(r['ref'])} | {c(r['gpu'])} | {c(r.get('env_blas'))} | {c(r.get('env_gelss'))} |")
                print("\n(χ²(α) − ν is a noisy, non-monotone function past the decade where the eps-truncation of the spectrum sets in; "
                      "when it crosses zero more than once the first crossing on the way down decides, and a rounding-level "
                      "difference can move it by several decades.  The fitted densities at either λ reproduce χ² = ν.)\n")


if __name__ == "__main__":
    main()
