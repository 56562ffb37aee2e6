"""TEST INFRASTRUCTURE -- tier-3 parity protocol of SURVEY.md 8-c item 3 (default / high order, where the
reference's coefficients are not reproducible to 1e-9 even against itself): per record, compare the GPU fit with
the oracle on

    * record status (NaN record / lambda = 0 / root found)            interpolate.py:188-191, 210-211, 558-563
    * scale factor and bracket decade of the chi2 = nu walk           interpolate.py:173-207
    * |log10 lambda_gpu - log10 lambda_ref|                           interpolate.py:214-216
    * numerical rank of X(lambda) (gelsd, rcond = eps)                interpolate.py:462
    * fitted densities A.C (what Estimate returns at the gates)       interpolate.py:566-569, estimate.py:115

and report the same quantities for the REFERENCE AGAINST ITSELF with its normal equations formed in BLAS order
instead of the einsum of interpolate.py:456 (a 1e-16 relative change of X): the reproducibility envelope.
Only tests/, bench.py's cpu_baseline leg and tools/ import this module; the product never does.
"""
import numpy as np

import ref_port as rp

ST_OK, ST_TOO_SMOOTH, ST_NO_ROOT = 0, 1, 2


def oracle_record(A_all, value_r, error_r, omega, name='curvature', order='einsum'):
    """Reference algorithm on one record.  order='einsum': the reference's own normal equations (bit for bit the
    reference's fit); 'blas': same algorithm, A^T W A formed by BLAS -- the envelope run; 'gelss': the reference's
    normal equations, lstsq through LAPACK gelss instead of gelsd (same rcond) -- the second envelope."""
    ok = np.isfinite(value_r)
    A = np.asfortranarray(A_all[ok])
    b = value_r[ok]
    W = np.array(error_r[ok] ** (-2))
    if order in ('einsum', 'gelss'):
        G, y = rp.normal_equations(A, W, b)
    else:
        AW = A * W[:, None]
        G, y = AW.T @ A, AW.T @ b
    out = rp.fit_record_given_normal_equations(A, b, W, G, y, {name: omega}, [name],
                                               lapack_driver='gelss' if order == 'gelss' else None)
    lam = out['lam'][name]
    out['status'] = ST_NO_ROOT if np.isnan(lam) else (ST_TOO_SMOOTH if lam == 0 else ST_OK)
    out['lam'] = float(lam)
    out['npts'] = int(ok.sum())
    out['AC'] = A @ out['C'] if np.isfinite(lam) else None
    return out


def _cmp(a, b, A_all, value_r):
    """a, b: dicts with status, sf, k_lo, lam, rank, C (b is the reference)."""
    d = {'status_same': bool(a['status'] == b['status'])}
    if a['status'] == ST_OK and b['status'] == ST_OK:
        d['sf_same'] = bool(abs(a['sf'] - b['sf']) < 1e-9)
        d['k_lo_same'] = bool(a['k_lo'] == b['k_lo'])
        d['dlog10_lambda'] = float(abs(np.log10(a['lam']) - np.log10(b['lam'])))
    if a['status'] in (ST_OK, ST_TOO_SMOOTH) and b['status'] == a['status']:
        ok = np.isfinite(value_r)
        da, db = A_all[ok] @ a['C'], A_all[ok] @ b['C']
        d['rank_diff'] = int(a['rank'] - b['rank'])
        d['AC_rel'] = float(np.max(np.abs(da - db)) / np.max(np.abs(db)))
        d['C_rel'] = float(np.max(np.abs(a['C'] - b['C'])) / np.max(np.abs(b['C'])))
        d['chi2_rel'] = float(abs(a['chi2'] - b['chi2']) / abs(b['chi2']))
    return d


def table_agreement(tab_gpu, tab_ref):
    """max relative difference of chi2(10^-k) over the decades both sides evaluated."""
    both = np.isfinite(tab_gpu) & np.isfinite(tab_ref)
    if not both.any():
        return None, 0
    return float(np.max(np.abs(tab_gpu[both] - tab_ref[both]) / np.abs(tab_ref[both]))), int(both.sum())


def gpu_record(res, r, trace=None):
    """Row r of a FitResult (numpy) in the dict form _cmp expects."""
    npts = None
    out = {'status': int(res.status[r]), 'lam': float(res.reg_params[r, 0]), 'rank': int(res.rank[r]),
           'C': np.asarray(res.Coeffs[r]), 'chi2': float(res.chi_sq[r]), 'sf': None, 'k_lo': None}
    if trace is not None:
        out['k_lo'] = int(trace['k_lo'][r])
        out['nu'] = float(trace['nu'][r])
        out['table'] = np.asarray(trace['table'][r])
    return out


def summarize(rows):
    """rows: list of _cmp dicts -> aggregate block for the bench line / profiles."""
    n = len(rows)
    def frac(key):
        v = [r[key] for r in rows if key in r]
        return {'agree': int(sum(v)), 'of': len(v)}
    def stats(key):
        v = np.array([r[key] for r in rows if key in r], dtype=float)
        if v.size == 0:
            return None
        return {'median': float(np.median(v)), 'max': float(v.max()), 'n': int(v.size)}
    return {'records': n, 'status_identical': frac('status_same'), 'scale_factor_identical': frac('sf_same'),
            'bracket_decade_identical': frac('k_lo_same'), 'abs_dlog10_lambda': stats('dlog10_lambda'),
            'rank_diff_abs': stats('rank_diff_abs'), 'AC_rel_diff': stats('AC_rel'), 'C_rel_diff': stats('C_rel'),
            'chi2_rel_diff': stats('chi2_rel'), 'table_rel_diff': stats('table_rel')}


def _pair(a, b, A_all, value_r):
    d = _cmp(a, b, A_all, value_r)
    if 'rank_diff' in d:
        d['rank_diff_abs'] = abs(d['rank_diff'])
    if a.get('table') is not None and b.get('table') is not None:
        t, nt = table_agreement(a['table'], b['table'])
        if t is not None:
            d['table_rel'] = t
            d['table_decades_compared'] = nt
    return d


def compare(gpu_rows, ref_rows, env_rows, A_all, value, alt_rows=None):
    """gpu_rows / ref_rows / env_rows / alt_rows: per-record dicts (same records, same order).  ref = the reference's
    algorithm as shipped (einsum normal equations, gelsd); env = the same with BLAS-order normal equations; alt = the
    same with LAPACK gelss instead of gelsd (both envelopes: equally valid executions of the reference)."""
    keys = ('status', 'sf', 'k_lo', 'lam', 'rank', 'calls')
    g, e, al, ga, per = [], [], [], [], []
    for r, (a, b, c) in enumerate(zip(gpu_rows, ref_rows, env_rows)):
        if a.get('sf') is None and a.get('nu') is not None and b.get('npts'):
            a['sf'] = a['nu'] / b['npts']
        dg = _pair(a, b, A_all, value[r])
        de = _pair(c, b, A_all, value[r]) if c is not None else {}
        row = {'record': r, 'ref': {k: b.get(k) for k in keys}, 'gpu': {k: a.get(k) for k in keys},
               'env_blas': {k: c.get(k) for k in keys} if c is not None else None,
               'gpu_vs_ref': dg, 'env_blas_vs_ref': de}
        g.append(dg); e.append(de)
        if alt_rows is not None:
            x = alt_rows[r]
            da, dga = _pair(x, b, A_all, value[r]), _pair(a, x, A_all, value[r])
            al.append(da); ga.append(dga)
            row.update(env_gelss={k: x.get(k) for k in keys}, env_gelss_vs_ref=da, gpu_vs_env_gelss=dga)
        per.append(row)
    out = {'gpu_vs_reference': summarize(g), 'reference_vs_itself_blas_order': summarize(e), 'per_record': per}
    if alt_rows is not None:
        out['reference_vs_itself_gelss_driver'] = summarize(al)
        out['gpu_vs_reference_with_gelss_driver'] = summarize(ga)
    return out
