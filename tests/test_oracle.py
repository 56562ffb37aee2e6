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
