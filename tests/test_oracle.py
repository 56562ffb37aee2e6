"""The oracle (oracle/ref_port.py) against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py).  The reference has no tests or fixtures of its own, so these
fixtures are what pins the oracle.  CPU only."""
import numpy as np
import pytest
import scipy.special as sp

from conftest import load_golden, oracle_model
import ref_port as rp

CASES = ["lo8", "lo12", "lo12_two", "mid27", "rbf27", "lo8_gcv"]


@pytest.mark.parametrize("name", CASES + ["c1_144"])
def test_basis_bit_identical(name):
    g = load_golden(name)
    A = oracle_model(g).basis(g["lat"], g["lon"], g["alt"])
    assert np.array_equal(A, g["A"])


@pytest.mark.parametrize("case", ["g12", "g144"])
def test_grad_basis_bit_identical(case):
    """oracle grad_basis == the unmodified reference's (fixture from oracle/make_golden_grad.py)."""
    import json
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "grad_basis.npz"))
    mk = json.loads(str(g[case + "_model_keys"]))
    m = rp.SphHarmLag(mk["MAXK"], mk["MAXL"], mk["CAP_LIM"], 78, 262)
    assert np.array_equal(m.grad_basis(g["lat"], g["lon"], g["alt"]), g[case + "_grad"])
    # and the value fixture of the same points agrees with the basis (ties the two fixtures together)
    assert np.array_equal(m.basis(g["lat"], g["lon"], g["alt"]), g[case + "_A"])


@pytest.mark.parametrize("name", CASES)
def test_fit_bit_identical(name):
    g = load_golden(name)
    m = oracle_model(g)
    regs = dict(zip(g["reglist"], g["regs"]))
    C, dC, c2, lam = rp.fit_records(m, g["lat"], g["lon"], g["alt"], g["value"], g["error"], regs, g["reglist"],
                                    method=g["default_keys"].get("REGULARIZATION_METHOD", "chi2"))
    assert np.array_equal(C, g["Coeffs"], equal_nan=True)
    assert np.array_equal(c2, g["chi_sq"], equal_nan=True)
    assert np.array_equal(dC, g["Covariance"], equal_nan=True)
    if g["reglist"]:
        assert np.array_equal(lam, g["lam"], equal_nan=True)


def test_fit_default_order_one_record():
    """N = 144 (example_config.ini order): one record of the C1-shaped fixture, bit-identical."""
    g = load_golden("c1_144")
    m = oracle_model(g)
    regs = dict(zip(g["reglist"], g["regs"]))
    C, dC, c2, lam = rp.fit_record(m, g["lat"], g["lon"], g["alt"], g["value"][0], g["error"][0], regs, g["reglist"],
                                   A_all=g["A"])
    assert np.array_equal(C, g["Coeffs"][0])
    assert c2 == g["chi_sq"][0]
    assert lam["curvature"] == g["lam"][0, 0]
    assert np.array_equal(np.diag(dC), g["Covariance_diag"][0])


@pytest.mark.parametrize("name", ["lo8", "lo12"])
def test_regularisation_matrices(name):
    g = load_golden(name)
    m = oracle_model(g)
    assert np.array_equal(m.omega(), g["reg_curvature"])


def test_psi_matrix():
    g = load_golden("lo12_two")
    assert np.array_equal(oracle_model(g).psi(), g["reg_0thorder"])


@pytest.mark.parametrize("name", CASES)
def test_estimate_and_hull(name):
    g = load_golden(name)
    m = oracle_model(g)
    assert np.array_equal(rp.hull_vertices(g["lat"], g["lon"], g["alt"]), g["hull_vert"])
    C = rp.select_coeffs(g["utime"], g["Coeffs"], float(g["q_time"]))
    assert np.array_equal(C, g["Coeffs"][int(g["q_record"])])
    out = rp.estimate(m, C, g["hull_vert"], g["q_lat"], g["q_lon"], g["q_alt"])
    assert np.array_equal(out, g["q_out"], equal_nan=True)
    # the O(F) half-space test takes the same decisions as the per-point re-hull of the reference
    eq = rp.hull_halfspaces(g["hull_vert"])
    x, y, z = rp.geodetic2ecef(g["q_lat"], g["q_lon"], g["q_alt"])
    inside = np.all(eq[:, 0, None, None] * x + eq[:, 1, None, None] * y + eq[:, 2, None, None] * z
                    + eq[:, 3, None, None] <= 0, axis=0)
    assert np.array_equal(inside, np.isfinite(g["q_out"]))
    assert inside.any() and (~inside).any()


def test_time_selection_errors():
    g = load_golden("lo8")
    with pytest.raises(ValueError):
        rp.select_coeffs(g["utime"], g["Coeffs"], float(g["utime"][-1, 1]) + 1000.0)
    Ci = rp.select_coeffs(g["utime"], g["Coeffs"], float(g["utime"][2].mean() + 10.0), timeinterp=True)
    mt = g["utime"].mean(axis=1)
    T = 10.0 / (mt[3] - mt[2])
    assert np.allclose(Ci, (1 - T) * g["Coeffs"][2] + T * g["Coeffs"][3], rtol=0, atol=0, equal_nan=True)


def test_lpmv_series_tracks_scipy():
    """Zhang & Jin restatement (what the CUDA kernel implements) vs the scipy binary."""
    rng = np.random.default_rng(0)
    worst = 0.0
    for cap in (6.0, 10.0, 15.0, 30.0):
        for l in range(0, 8):
            v = (2 * l + 0.5) * np.pi / (2 * np.radians(cap)) - 0.5
            for m in range(-l, l + 1):
                for th in rng.uniform(0.5, 45.0, 4):
                    x = np.cos(np.radians(th))
                    a, b = rp.lpmv_series(m, v, x), sp.lpmv(m, v, x)
                    if b != 0 and abs(b) > 1e-290:
                        worst = max(worst, abs(a - b) / max(abs(b), 1e-300))
    assert worst < 1e-11


def test_einsum_is_sequential_two_rounding_sum():
    """The property the strict CUDA kernel relies on: np.einsum('ji,j,jk->ik') (interpolate.py:456)
    equals acc = (A[j,i]*W[j])*A[j,k] + acc summed over j in order, without fused multiply-add."""
    g = load_golden("lo12")
    ok = np.isfinite(g["value"][0])
    A, W, b = g["A"][ok], g["error"][0][ok] ** -2, g["value"][0][ok]
    G, y = rp.normal_equations(A, W, b)
    acc = np.zeros_like(G)
    yy = np.zeros_like(y)
    for j in range(A.shape[0]):
        t = A[j] * W[j]
        acc = np.outer(t, A[j]) + acc
        yy = t * b[j] + yy
    assert np.array_equal(acc, G)
    assert np.array_equal(yy, y)
    # zero-weight masking == row deletion (bit for bit)
    Wm = np.where(ok, np.nan_to_num(g["error"][0]) ** -2 if False else 0.0, 0.0)
    Wm[ok] = W
    bm = np.where(ok, g["value"][0], 0.0)
    G2, y2 = rp.normal_equations(g["A"], Wm, bm)
    assert np.array_equal(G2, G) and np.array_equal(y2, y)


def test_search_given_normal_equations_is_the_reference_bit_for_bit():
    """ref_port.fit_record_given_normal_equations (the fast form the parity protocol and the envelope fixtures use:
    A^T W A formed once instead of at every trial) == the unmodified reference: same lambda, same coefficients, and
    the same sequence of (alpha, chi2 - nu) evaluations as the golden trace."""
    import parity
    for name in ("lo12", "mid27"):
        g = load_golden(name)
        for r in range(g["value"].shape[0]):
            o = parity.oracle_record(g["A"], g["value"][r], g["error"][r], g["regs"][0], g["reglist"][0])
            assert np.array_equal(o["C"], g["Coeffs"][r], equal_nan=True)
            assert (np.isnan(o["lam"]) and np.isnan(g["lam"][r, 0])) or o["lam"] == g["lam"][r, 0]
            if "trace" in g:
                tr = g["trace"][r]
                tr = tr[np.isfinite(tr[:, 0])]
                assert o["calls"] == len(tr)
                tab = rp._decade_table([tuple(x) for x in tr])
                both = np.isfinite(tab) & np.isfinite(o["table"])
                assert both.sum() >= 1 and np.allclose(tab[both], o["table"][both], rtol=1e-14, atol=0)


def test_envelope_fixture_matches_the_golden_reference_run():
    """tests/golden/envelope_*.json: the 'gelsd' row is the reference itself (lambda equal to the golden run's); the
    three executions agree on scale factor and bracket decade, and genuinely disagree on lambda at N = 144."""
    import json
    import os
    from conftest import GOLDEN
    for name in ("mid27", "c1_144", "c3_500"):
        g = load_golden(name)
        env = json.load(open(os.path.join(GOLDEN, f"envelope_{name}.json")))["records"]
        assert len(env) == g["value"].shape[0]
        for r, e in enumerate(env):
            lam = g["lam"][r, 0]
            if np.isnan(lam):
                assert e["gelsd"]["status"] == 2
                continue
            assert e["gelsd"]["lam"] == lam
            assert len({e[d]["sf"] for d in ("gelsd", "blas", "gelss")}) == 1
            assert len({e[d]["k_lo"] for d in ("gelsd", "blas", "gelss")}) == 1
    env = json.load(open(os.path.join(GOLDEN, "envelope_c1_144.json")))["records"]
    spread = [abs(np.log10(e["gelsd"]["lam"]) - np.log10(e["gelss"]["lam"])) for e in env]
    assert max(spread) > 0.05        # the reference does not reproduce its own lambda across LAPACK drivers


def test_hull_tolerance_matches_the_rehull_decision():
    """estimate.hull_halfspaces (half-spaces with Qhull's roundoff allowance folded in) takes the reference's
    decision (estimate.py:167-177) for points ON the facets of the hull and for points clearly inside / outside.
    A query point that DUPLICATES a hull vertex is the one documented deviation: the reference's answer there depends
    on which of the two coincident points Qhull keeps as the vertex (its vertex list then differs from the saved
    one and the point counts as outside: 3 of the 18 vertices of this hull, 15 count as inside); the half-space
    test says inside for all of them."""
    from scipy.spatial import ConvexHull
    from volumetricinterp_b200.estimate import hull_halfspaces
    g = load_golden("c1_144")
    hv = g["hull_vert"]
    eq = hull_halfspaces(hv)
    inside = lambda p: bool(np.all(eq[:, :3] @ p + eq[:, 3] <= 0.0))
    base = ConvexHull(hv).vertices
    rehull = lambda p: bool(np.array_equal(base, ConvexHull(np.vstack([hv, p[None]])).vertices))
    hull = ConvexHull(hv)
    cent = hv[hull.simplices].mean(axis=1)
    pts = [c for c in cent]
    pts += [c + 1.0 * n for c, n in zip(cent, hull.equations[:, :3])]        # 1 m outside
    pts += [c - 1.0 * n for c, n in zip(cent, hull.equations[:, :3])]        # 1 m inside
    for p in pts:
        assert inside(p) == rehull(p)
    at_vertices = [rehull(v) for v in hv]
    assert all(inside(v) for v in hv)
    assert 0 < sum(at_vertices) < len(hv)        # the reference itself is not consistent on its own vertices


def test_high_order_golden_is_reproduced_by_the_oracle():
    """c3_500 (N = 500: MAXK 5, MAXL 10, CAP_LIM 11; generated by the unmodified reference with its real curvature
    matrix): the oracle's design matrix is bit-identical and its search lands on the golden lambda and coefficients."""
    import parity
    g = load_golden("c3_500")
    m = oracle_model(g)
    assert m.nbasis == 500
    assert np.array_equal(m.basis(g["lat"], g["lon"], g["alt"]), g["A"])
    r = 1                                                   # 53 eval_C calls in the reference's trace
    o = parity.oracle_record(g["A"], g["value"][r], g["error"][r], g["regs"][0], g["reglist"][0])
    assert o["lam"] == g["lam"][r, 0] and o["calls"] == int(g["n_eval"][r])
    assert np.array_equal(o["C"], g["Coeffs"][r])
