"""Device-agnostic scalar code of volumetricinterp_b200/csrc (vi_math.h, vi_tridiag.h, vi_tql.h,
vi_brent.h) compiled for the CPU by the TEST-ONLY harness and checked against the golden vectors,
scipy and LAPACK.  The product never loads this harness; it has no CPU path."""
import ctypes as C
import io
import os

import numpy as np
import pytest
import scipy.linalg
import scipy.optimize

from conftest import GOLDEN, dptr, load_golden, product_model
import ref_port as rp

EPS = np.finfo(float).eps


def _rel(a, b):
    scale = np.maximum(np.abs(b), 1e-12 * np.abs(b).max())
    return np.max(np.abs(a - b) / scale)


@pytest.mark.parametrize("name", ["lo8", "lo12", "mid27", "c1_144"])
def test_sphharmlag_rows_match_reference_basis(harness, name):
    g = load_golden(name)
    m = product_model(g)
    P = m.params()
    assert harness.h_sizeof_shl_params() == C.sizeof(P)
    n = g["lat"].size
    A = np.zeros((n, m.nbasis))
    harness.h_shl_rows(C.byref(P), dptr(g["lat"]), dptr(g["lon"]), dptr(g["alt"]), C.c_int64(n), dptr(A))
    # column-relative tolerance: entries near a zero of P_v^m are compared against the column scale
    for c in range(m.nbasis):
        ref = g["A"][:, c]
        tol = 2e-12 * max(np.abs(ref).max(), 1e-300)
        assert np.max(np.abs(A[:, c] - ref)) <= tol, (c, np.max(np.abs(A[:, c] - ref)), np.abs(ref).max())


@pytest.mark.parametrize("case", ["g12", "g144"])
def test_sphharmlag_gradient_rows_match_reference(harness, case):
    """vi_shl_grad_row (csrc/vi_math.h) against the reference's own grad_basis (sphharmlag.py:148-184; fixture
    written by oracle/make_golden_grad.py from the unmodified reference): component- and column-relative 2e-12."""
    import io
    from volumetricinterp_b200.models import sphharmlag
    g = np.load(os.path.join(GOLDEN, "grad_basis.npz"))
    m = sphharmlag.Model(io.StringIO(str(g[case + "_config_text"])))
    P = m.params()
    n = g["lat"].size
    out = np.zeros((n, 3, m.nbasis))
    harness.h_shl_grad_rows(C.byref(P), dptr(g["lat"]), dptr(g["lon"]), dptr(g["alt"]), C.c_int64(n), dptr(out))
    ref = g[case + "_grad"]
    assert out.shape == ref.shape
    for comp in range(3):
        for c in range(m.nbasis):
            r = ref[:, comp, c]
            tol = 2e-12 * max(np.abs(r).max(), 1e-300)
            assert np.max(np.abs(out[:, comp, c] - r)) <= tol, (comp, c, np.max(np.abs(out[:, comp, c] - r)), np.abs(r).max())


def test_radbasfun_rows_match_reference_basis(harness):
    g = load_golden("rbf27")
    m = product_model(g)
    n = g["lat"].size
    A = np.zeros((n, m.nbasis))
    harness.h_rbf_rows(dptr(g["lat"]), dptr(g["lon"]), dptr(g["alt"]), C.c_int64(n), dptr(m.centers),
                       m.nbasis, C.c_double(m.eps), dptr(A))
    assert np.allclose(A, g["A"], rtol=1e-12, atol=1e-300)


PACKED = -1
TWO_STAGE = -2      # band reduction + bulge chasing (vi_band.h, vi_chase.h), device code run by tests/cuda_emu.h


def _system(harness, G, y, regs, lam, nt):
    n = G.shape[0]
    Cq = np.zeros(n)
    rank, bad = C.c_int(0), C.c_int(0)
    dd, ee = np.zeros(n), np.zeros(n)
    regs = np.ascontiguousarray(regs)
    lam = np.ascontiguousarray(lam, dtype=float)
    if nt == TWO_STAGE:
        st = harness.h_system_solve_two_stage(n, dptr(G), dptr(y), dptr(regs), dptr(lam), len(lam), C.c_double(EPS),
                                              dptr(Cq), C.byref(rank), dptr(dd), dptr(ee), C.byref(bad), None)
    elif nt == PACKED:      # the packed-triangle kernel's phases (vi_tridiag_packed.h), thread count fixed by n
        st = harness.h_system_solve_packed(n, dptr(G), dptr(y), dptr(regs), dptr(lam), len(lam), C.c_double(EPS),
                                           dptr(Cq), C.byref(rank), dptr(dd), dptr(ee), C.byref(bad))
    else:
        st = harness.h_system_solve(n, dptr(G), dptr(y), dptr(regs), dptr(lam), len(lam), nt, C.c_double(EPS),
                                    dptr(Cq), C.byref(rank), dptr(dd), dptr(ee), C.byref(bad))
    return st, bad.value, rank.value, Cq, dd, ee


@pytest.mark.parametrize("name,nt", [("lo8", 8), ("lo8", 32), ("lo12", 48), ("lo12_two", 12),
                                     ("lo8", PACKED), ("lo12", PACKED), ("lo12_two", PACKED),
                                     ("lo8", TWO_STAGE), ("lo12", TWO_STAGE), ("lo12_two", TWO_STAGE)])
def test_system_pipeline_matches_lstsq_low_order(harness, name, nt):
    """tridiagonalise (CTA phases run thread by thread) + tape QL + truncated solve + back-transform
    == scipy.linalg.lstsq (interpolate.py:462) on full-rank systems."""
    g = load_golden(name)
    for r in range(g["value"].shape[0]):
        ok = np.isfinite(g["value"][r])
        A, W, b = g["A"][ok], g["error"][r][ok] ** -2, g["value"][r][ok]
        G, y = rp.normal_equations(A, W, b)
        for alpha in (0.0, -12.0, -24.0, -30.0):
            lam = np.zeros(len(g["regs"]))
            lam[0] = 10.0 ** alpha
            X = G + sum(l * R for l, R in zip(lam, g["regs"]))
            ref = scipy.linalg.lstsq(X, y)[0]
            st, bad, rank, Cq, dd, ee = _system(harness, np.ascontiguousarray(G), np.ascontiguousarray(y),
                                                np.stack(g["regs"]), lam, nt)
            assert st == 0 and bad == 0 and rank == G.shape[0]
            # tridiagonal form has the same spectrum
            T = np.diag(dd) + np.diag(ee[:-1], 1) + np.diag(ee[:-1], -1)
            ev = np.linalg.eigvalsh(0.5 * (X + X.T))
            scl = np.abs(ev).max() / np.abs(np.linalg.eigvalsh(T)).max()
            assert np.allclose(np.linalg.eigvalsh(T) * scl, ev, rtol=0, atol=1e-13 * np.abs(ev).max())
            cond = np.abs(ev).max() / np.abs(ev).min()
            assert np.max(np.abs(Cq - ref)) <= 50 * EPS * cond * np.abs(ref).max()


@pytest.mark.parametrize("nt", [576, PACKED, TWO_STAGE])
def test_system_pipeline_rank_deficient(harness, nt):
    """N = 144: numerical rank and fitted densities agree with lstsq where the solve is well posed."""
    g = load_golden("c1_144")
    ok = np.isfinite(g["value"][0])
    A, W, b = g["A"][ok], g["error"][0][ok] ** -2, g["value"][0][ok]
    G, y = rp.normal_equations(A, W, b)
    for alpha in (0.0, -10.0):
        X = G + 10.0 ** alpha * g["regs"][0]
        ref = scipy.linalg.lstsq(X, y)[0]
        s = np.linalg.svd(X, compute_uv=False)
        st, bad, rank, Cq, _, _ = _system(harness, np.ascontiguousarray(G), np.ascontiguousarray(y),
                                          np.stack(g["regs"]), [10.0 ** alpha], nt)
        assert st == 0 and bad == 0
        assert abs(rank - int((s > EPS * s[0]).sum())) <= 1
        dens, dref = A @ Cq, A @ ref
        assert np.max(np.abs(dens - dref)) <= 1e-5 * np.abs(dref).max()


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 16, 27, 40, 65, 100, 144, 150, 200])
def test_packed_tridiagonalisation_random(harness, n):
    """Packed-triangle Householder reduction: T has the spectrum of X, the solve matches numpy, for orders
    around every octet / warp-assignment boundary."""
    rng = np.random.default_rng(100 + n)
    M = rng.standard_normal((n, n))
    X = M @ M.T + n * np.eye(n)
    y = rng.standard_normal(n)
    st, bad, rank, Cq, dd, ee = _system(harness, np.ascontiguousarray(X), y, np.zeros((1, n, n)), [0.0], PACKED)
    assert bad == 0
    T = np.diag(dd) + np.diag(ee[:-1], 1) + np.diag(ee[:-1], -1)
    ev = np.linalg.eigvalsh(X)
    evT = np.linalg.eigvalsh(T)
    scl = np.abs(ev).max() / np.abs(evT).max()
    assert np.allclose(evT * scl, ev, rtol=0, atol=1e-13 * np.abs(ev).max())
    assert st == 0 and rank == n     # (the rotation tape holds 2 n^2 + 64 entries; these spectra need ~1.4 n^2)
    ref = np.linalg.solve(X, y)
    assert np.max(np.abs(Cq - ref)) <= 1e-12 * np.abs(ref).max()
    # same system through the full-square phases: same tridiagonal form up to rounding
    st2, bad2, rank2, Cq2, dd2, ee2 = _system(harness, np.ascontiguousarray(X), y, np.zeros((1, n, n)), [0.0],
                                              max(32, (n + 31) // 32 * 32))
    assert np.allclose(dd, dd2, rtol=0, atol=1e-12 * np.abs(dd2).max())
    assert np.allclose(np.abs(ee), np.abs(ee2), rtol=0, atol=1e-12 * np.abs(dd2).max())


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 9, 15, 16, 17, 27, 40, 65, 100, 144, 168])
def test_two_stage_tridiagonalisation_random(harness, n):
    """Band reduction (DMMA block reflectors, one CTA) + bulge chasing (one warp): the device code of vi_band.h /
    vi_chase.h executed by the CUDA-on-CPU model.  T has the spectrum of X and Q1 Q2 T+ Q2^T Q1^T y solves the
    system, for orders around every block / warp-count boundary up to the largest the two-stage path takes."""
    rng = np.random.default_rng(300 + n)
    M = rng.standard_normal((n, n))
    X = M @ M.T + n * np.eye(n)
    y = rng.standard_normal(n)
    st, bad, rank, Cq, dd, ee = _system(harness, np.ascontiguousarray(X), y, np.zeros((1, n, n)), [0.0], TWO_STAGE)
    assert bad == 0 and st == 0 and rank == n
    T = np.diag(dd) + np.diag(ee[:-1], 1) + np.diag(ee[:-1], -1)
    ev, evT = np.linalg.eigvalsh(X), np.linalg.eigvalsh(T)
    scl = np.abs(ev).max() / np.abs(evT).max()
    assert np.allclose(evT * scl, ev, rtol=0, atol=1e-13 * np.abs(ev).max())
    ref = np.linalg.solve(X, y)
    assert np.max(np.abs(Cq - ref)) <= 1e-12 * np.abs(ref).max()


def test_two_stage_band_is_orthogonally_similar(harness):
    """Stage 1 alone: the band it hands to stage 2 (half-width 8) has the spectrum of X, also for an indefinite,
    graded matrix like the regularised systems of the fit."""
    n = 100
    rng = np.random.default_rng(5)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = rng.standard_normal(n) * 10.0 ** rng.uniform(-14, 0, n)
    X = (Q * ev) @ Q.T
    X = 0.5 * (X + X.T)
    y = rng.standard_normal(n)
    npad = (n + 7) // 8 * 8
    band = np.zeros(9 * npad)
    Cq, dd, ee = np.zeros(n), np.zeros(n), np.zeros(n)
    rank, bad = C.c_int(0), C.c_int(0)
    regs, lam = np.zeros((1, n, n)), np.zeros(1)
    st = harness.h_system_solve_two_stage(n, dptr(X), dptr(y), dptr(regs), dptr(lam), 1, C.c_double(EPS), dptr(Cq),
                                          C.byref(rank), dptr(dd), dptr(ee), C.byref(bad), dptr(band))
    assert st == 0 and bad.value == 0
    Bm = np.zeros((n, n))
    for j in range(n):
        for d in range(9):
            if j + d < n:
                Bm[j + d, j] = Bm[j, j + d] = band[9 * j + d]
    scl = 2.0 ** -np.frexp(np.abs(X).max())[1]
    evb = np.sort(np.linalg.eigvalsh(Bm) / scl)
    assert np.allclose(evb, np.sort(np.linalg.eigvalsh(X)), rtol=0, atol=1e-14 * np.abs(ev).max())
    assert np.abs(band.reshape(npad, 9)[n:]).max(initial=0.0) == 0.0


@pytest.mark.parametrize("n,G", [(144, 8), (144, 16), (27, 8), (97, 8)])
def test_wavefront_replay_several_systems_per_warp(harness, n, G):
    """vi_wav_pass<., G>: 32 / G systems of different spectra (hence different sweep structures and lengths) replayed
    side by side by the lane groups of one warp, each bit-identical to its sequential replay."""
    S = 32 // G
    rng = np.random.default_rng(n + G)
    d = rng.standard_normal((S, n))
    e = rng.standard_normal((S, n))
    for q in range(S):
        if q % 2 == 0:
            d[q] *= 10.0 ** rng.uniform(-12, 0, n)
            e[q] *= 10.0 ** rng.uniform(-12, 0, n)
        if q == 1:
            e[q, n // 2] = 0.0
    g = rng.standard_normal((S, n))
    fn = harness.h_wave_replay_groups
    fn.restype = C.c_int
    assert fn(n, G, dptr(np.ascontiguousarray(d)), dptr(np.ascontiguousarray(e)), dptr(np.ascontiguousarray(g))) == 0


def _band_of(harness, X, y, split):
    n = X.shape[0]
    npad = (n + 7) // 8 * 8
    band = np.zeros(9 * npad)
    Cq, dd, ee = np.zeros(n), np.zeros(n), np.zeros(n)
    rank, bad = C.c_int(0), C.c_int(0)
    regs, lam = np.zeros((1, n, n)), np.zeros(1)
    harness.h_bnd_set_split(*split)
    try:
        st = harness.h_system_solve_two_stage(n, dptr(X), dptr(y), dptr(regs), dptr(lam), 1, C.c_double(EPS), dptr(Cq),
                                              C.byref(rank), dptr(dd), dptr(ee), C.byref(bad), dptr(band))
    finally:
        harness.h_bnd_set_split(0, 0, 4)
    assert st == 0 and bad.value == 0
    return band, Cq, dd, ee, rank.value


@pytest.mark.parametrize("n,p1", [(144, 7), (144, 1), (100, 5), (27, 1), (168, 10)])
def test_two_stage_split_reduction(harness, n, p1):
    """k_band + k_band_tail: the panels [0, p1) in one kernel, the trailing matrix (order n - 8 p1, a quarter of the
    shared memory at n = 144: four CTAs per SM instead of two) in a second one.  With the same warp count and panel
    register rows the hand-over changes nothing: band, tridiagonal and solution are BIT-IDENTICAL to the unsplit
    reduction.  The production geometry of the tail (4 warps, 3 register rows per lane) only re-orders sums."""
    rng = np.random.default_rng(900 + n + p1)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = rng.standard_normal(n) * 10.0 ** rng.uniform(-10, 0, n)
    X = (Q * ev) @ Q.T
    X = np.ascontiguousarray(0.5 * (X + X.T))
    y = rng.standard_normal(n)
    b0, c0, d0, e0, r0 = _band_of(harness, X, y, (0, 0, 4))
    nw_full = harness.h_bnd_threads(n) // 32
    b1, c1, d1, e1, r1 = _band_of(harness, X, y, (p1, nw_full, 4))
    if harness.h_bnd_threads(n - 8 * p1) // 32 == nw_full:      # (a smaller trailing order may get fewer warps by default)
        pass
    assert np.array_equal(b0, b1) and np.array_equal(d0, d1) and np.array_equal(e0, e1) and np.array_equal(c0, c1)
    b2, c2, d2, e2, r2 = _band_of(harness, X, y, (p1, 4, 3))
    scale = np.abs(b0).max()
    assert np.allclose(b2, b0, rtol=0, atol=2e-14 * scale)
    T0 = np.diag(d0) + np.diag(e0[:-1], 1) + np.diag(e0[:-1], -1)
    T2 = np.diag(d2) + np.diag(e2[:-1], 1) + np.diag(e2[:-1], -1)
    assert np.allclose(np.linalg.eigvalsh(T2), np.linalg.eigvalsh(T0), rtol=0, atol=1e-13 * np.abs(d0).max())


@pytest.mark.parametrize("n", [9, 40, 144, 200])
def test_two_stage_global_memory_form(harness, n):
    """k_band_big (orders beyond 168: X in global memory, panel QR on the panel in shared memory, no look-ahead): same
    reduction as the shared-memory kernel up to summation order, also on a graded indefinite matrix."""
    rng = np.random.default_rng(700 + n)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = rng.standard_normal(n) * 10.0 ** rng.uniform(-10, 0, n)
    X = (Q * ev) @ Q.T
    X = np.ascontiguousarray(0.5 * (X + X.T))
    y = rng.standard_normal(n)
    harness.h_bnd_set_big(1)
    try:
        b1, c1, d1, e1, r1 = _band_of(harness, X, y, (0, 0, 4))
    finally:
        harness.h_bnd_set_big(0)
    npad = (n + 7) // 8 * 8
    Bm = np.zeros((n, n))
    for j in range(n):
        for d in range(9):
            if j + d < n:
                Bm[j + d, j] = Bm[j, j + d] = b1[9 * j + d]
    scl = 2.0 ** -np.frexp(np.abs(X).max())[1]
    assert np.allclose(np.sort(np.linalg.eigvalsh(Bm) / scl), np.sort(np.linalg.eigvalsh(X)), rtol=0, atol=1e-13 * np.abs(ev).max())
    T1 = np.diag(d1) + np.diag(e1[:-1], 1) + np.diag(e1[:-1], -1)
    assert np.allclose(np.sort(np.linalg.eigvalsh(T1) / scl), np.sort(np.linalg.eigvalsh(X)), rtol=0, atol=1e-13 * np.abs(ev).max())
    if n <= 168:
        b0, c0, d0, e0, r0 = _band_of(harness, X, y, (0, 0, 4))
        assert np.allclose(np.abs(b1), np.abs(b0), rtol=0, atol=1e-13 * np.abs(b0).max())
    # and the solve through it on a well-conditioned system
    M = rng.standard_normal((n, n))
    Xw = np.ascontiguousarray(M @ M.T + n * np.eye(n))
    harness.h_bnd_set_big(1)
    try:
        st, bad, rank, Cq, dd, ee = _system(harness, Xw, y, np.zeros((1, n, n)), [0.0], TWO_STAGE)
    finally:
        harness.h_bnd_set_big(0)
    assert st == 0 and bad == 0 and rank == n
    ref = np.linalg.solve(Xw, y)
    assert np.max(np.abs(Cq - ref)) <= 1e-12 * np.abs(ref).max()


def test_two_stage_layout(harness):
    """Block layout of vi_band.h: the element map of a block is a bijection, and each of the three fragment access
    patterns touches 16 distinct 8-byte bank pairs per half warp (64-bit accesses) / 8 distinct 16-byte slots per
    quarter warp (128-bit accumulator pairs): shared-memory bank-conflict free.  Reflector offsets of vi_chase.h
    match a direct count."""
    el = np.array([[harness.h_bnd_el(r, c) for c in range(8)] for r in range(8)])
    assert sorted(el.ravel()) == list(range(64))
    lanes = np.arange(32)
    for t in (0, 1):
        a_norm = np.array([el[l // 4, 4 * t + l % 4] for l in lanes])          # A operand of a stored block
        a_tran = np.array([el[4 * t + l % 4, l // 4] for l in lanes])          # A operand of its transpose
        for pat in (a_norm, a_tran):
            for half in (pat[:16], pat[16:]):
                assert len(set(half % 16)) == 16
    acc = np.array([el[l // 4, 2 * (l % 4)] for l in lanes])                     # accumulator pair (even column)
    assert (acc % 2 == 0).all() and all(el[l // 4, 2 * (l % 4) + 1] == acc[l] + 1 for l in lanes)
    for q in range(4):
        assert len(set((acc[8 * q:8 * q + 8] // 2) % 8)) == 8
    for n in (3, 9, 16, 17, 100, 144, 168):
        off = 0
        for s in range(max(n - 2, 0)):
            assert harness.h_chs_off(n, s) == off
            off += (n - 1 - s + 7) // 8
        assert harness.h_chs_nrefl(n) == off
    assert harness.h_bnd_threads(144) == 192 and harness.h_bnd_smem_bytes(144) <= (233472 - 2048) // 2      # two CTAs per SM


@pytest.mark.parametrize("N", [1, 8, 12, 16, 27, 48, 100, 144, 150, 160])
def test_normal_equation_work_split(harness, N):
    """k_ne_dmma3's units (lower-triangle tiles + one diagonal unit per row block) are dealt to the 16 warps exactly
    once, at most 7 per warp, contiguously in (row block, tile) order, with balanced tensor sub-partitions."""
    W = harness.h_ne3_warps()
    uinfo = (C.c_int * 96)()
    wbeg, wend = (C.c_int * W)(), (C.c_int * W)()
    nunits = C.c_int(0)
    ok = harness.h_ne3_split(N, 7, C.byref(nunits), uinfo, wbeg, wend)
    mt = (N + 15) // 16
    if mt * (mt + 1) > 96:          # does not fit the unit table: the launcher falls back to version 2
        assert ok == 0
        return
    assert ok == 1 and nunits.value == mt * (mt + 1)
    # every lower-triangle element is covered by exactly one unit: full tiles cover rows 16 mi.., columns 8 ni..;
    # diagonal units the lower 8 rows of tile (mi, 2 mi + 1)
    cover = np.zeros((16 * mt, 16 * mt), dtype=int)
    seen = np.zeros(nunits.value, dtype=int)
    weights = np.zeros(W)
    for w in range(W):
        assert 0 <= wbeg[w] <= wend[w] <= nunits.value and wend[w] - wbeg[w] <= 7
        for u in range(wbeg[w], wend[w]):
            seen[u] += 1
            mi, ni, diag, half = uinfo[u] & 255, (uinfo[u] >> 8) & 255, bool(uinfo[u] & (1 << 16)), bool(uinfo[u] & (1 << 17))
            weights[w] += 3 if diag else 4
            if not diag:
                cover[16 * mi:16 * mi + 16, 8 * ni:8 * ni + 8] += 1
            else:
                assert ni == 2 * mi + 1 and half == (8 * ni < N)
                if half:
                    cover[16 * mi + 8:16 * mi + 16, 8 * ni:8 * ni + 8] += 1
    assert (seen == 1).all()
    ii, kk = np.tril_indices(N)
    assert (cover[ii, kk] == 1).all()
    sp = [weights[p::4].sum() for p in range(4)]
    assert max(sp) - min(sp) <= 8


@pytest.mark.parametrize("n", [1, 7, 8, 9, 27, 64, 100, 144, 150, 208])
def test_packed_layout_and_octet_assignment(harness, n):
    """Storage map of the packed tridiagonalisation: every stored position (i >= 8 floor(c / 8)) has its own slot and
    the slots fill the array exactly; tile slots likewise; and at every step each active octet belongs to exactly
    one (warp, slot)."""
    npad, noct, nwarp, xd, nt = (C.c_int(0) for _ in range(5))
    total = harness.h_trp_geometry(n, *(C.byref(v) for v in (npad, noct, nwarp, xd, nt)))
    npad, noct, nwarp, xd, nt = (v.value for v in (npad, noct, nwarp, xd, nt))
    assert npad % 8 == 0 and npad >= n and noct == npad // 8 and total >= xd + 2 * nt
    idx = [harness.h_trp_idx(n, i, c) for c in range(npad) for i in range(8 * (c // 8), npad)]
    assert sorted(idx) == list(range(xd))
    # the two rows of a row pair are adjacent (one 128-bit access) and columns are contiguous in i
    assert all(harness.h_trp_idx(n, i + 1, c) == harness.h_trp_idx(n, i, c) + 1
               for c in range(0, npad, 5) for i in range(8 * (c // 8), npad - 1))
    slots = [harness.h_trp_tile_slot(n, ip, q) for q in range(noct) for ip in range(4 * q, npad // 2)]
    assert sorted(slots) == list(range(nt))
    for a in range(noct):
        got = [harness.h_trp_octet_of(n, w, sl, a) for w in range(nwarp) for sl in (0, 1)]
        got = sorted(q for q in got if q >= 0)
        assert got == list(range(a, noct))


def test_nonfinite_system_is_flagged(harness):
    G = np.eye(4)
    G[1, 2] = np.inf
    st, bad, *_ = _system(harness, G, np.ones(4), np.zeros((1, 4, 4)), [0.0], 4)
    assert bad == 1


def test_tql_against_lapack(harness):
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 17, 64):
        d = rng.standard_normal(n) * 10.0 ** rng.uniform(-8, 0, n)
        e = rng.standard_normal(max(n - 1, 0)) * 10.0 ** rng.uniform(-8, 0, max(n - 1, 0))
        g0 = rng.standard_normal(n)
        T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
        w = np.zeros(n); lam = np.zeros(n)
        rank, nrot = C.c_int(0), C.c_int(0)
        st = harness.h_tql_solve(n, dptr(d), dptr(np.append(e, 0.0)), dptr(g0), C.c_double(EPS), dptr(w), dptr(lam),
                                 C.byref(rank), C.byref(nrot))
        assert st == 0
        ev, V = np.linalg.eigh(T)
        assert np.allclose(np.sort(lam), ev, rtol=0, atol=4 * EPS * np.abs(ev).max() * n)
        keep = np.abs(ev) > EPS * np.abs(ev).max()
        ref = V[:, keep] @ ((V[:, keep].T @ g0) / ev[keep])
        assert rank.value == keep.sum()
        cond = np.abs(ev).max() / np.abs(ev[keep]).min()
        assert np.max(np.abs(w - ref)) <= 100 * EPS * cond * np.abs(ref).max()
        # the variant the GPU runs: eigenvalues + tape, right-hand side through tape replays
        w2 = np.zeros(n); lam2 = np.zeros(n)
        st = harness.h_tql_values_solve(n, dptr(d), dptr(np.append(e, 0.0)), dptr(g0), C.c_double(EPS), dptr(w2),
                                        dptr(lam2), C.byref(rank), C.byref(nrot))
        assert st == 0 and rank.value == keep.sum()
        assert np.allclose(np.sort(lam2), ev, rtol=0, atol=4 * EPS * np.abs(ev).max() * n)
        assert np.max(np.abs(w2 - ref)) <= 100 * EPS * cond * np.abs(ref).max()


def test_brentq_state_machine_replays_scipy(harness):
    """Same abscissae, same root as scipy.optimize.brentq (interpolate.py:214)."""
    FN = C.CFUNCTYPE(C.c_double, C.c_double)
    for f, a, b in [(lambda x: np.tanh(3 * (x + 24.3)) * 40 + 1.5, -25.0, -24.0),
                    (lambda x: (x + 7.123456789) * (1 + (x + 7) ** 2), -8.0, -7.0),
                    (lambda x: np.exp(x + 50.5) - 1.0, -51.0, -50.0)]:
        xs_ref = []
        def fr(x):
            xs_ref.append(x)
            return f(x)
        root_ref = scipy.optimize.brentq(fr, a, b, disp=True)
        xs = np.zeros(256)
        root, nfev = C.c_double(0), C.c_int(0)
        done = harness.h_brentq(FN(lambda x: float(f(x))), C.c_double(a), C.c_double(b), C.byref(root), dptr(xs),
                                C.byref(nfev))
        assert done == 1
        assert root.value == root_ref
        assert list(xs[:nfev.value]) == xs_ref
    # a triple root does not converge within maxiter=100: scipy raises, the state machine says done == 2
    f = lambda x: (x + 7.123456789) ** 3
    with pytest.raises(RuntimeError):
        scipy.optimize.brentq(f, -8.0, -7.0, disp=True)
    xs = np.zeros(256)
    root, nfev = C.c_double(0), C.c_int(0)
    assert harness.h_brentq(FN(lambda x: float(f(x))), C.c_double(-8.0), C.c_double(-7.0), C.byref(root), dptr(xs),
                            C.byref(nfev)) == 2


def _walk_python(table, npts):
    """interpolate.py:173-211 on a table of chi2(10^-k)."""
    bracket = False
    for sf in (0.6, 0.7, 0.8, 0.9, 1.0):
        nu = npts * sf
        k, val0, val = 0, 1.0, table[0] - nu
        if val < 0:
            return 1, -1, nu
        while val0 * val > 0:
            bracket = True
            val0 = val
            k += 1
            val = table[k] - nu
            if k > 100:
                bracket = False
                break
        if bracket:
            return 0, k, nu
    return 2, -1, 0.0


def test_bracket_walk(harness):
    rng = np.random.default_rng(7)
    for trial in range(200):
        npts = int(rng.integers(50, 800))
        kind = trial % 4
        k0 = rng.integers(1, 100)
        if kind == 0:      # monotone decrease through nu somewhere
            table = npts * (1.5 - 1.0 / (1 + np.exp(-(np.arange(102) - k0))))
        elif kind == 1:    # never below: no root
            table = npts * (1.2 + rng.uniform(0, 1, 102))
        elif kind == 2:    # too smooth at alpha = 0
            table = npts * rng.uniform(0.1, 0.5, 102)
        else:              # noisy
            table = npts * rng.uniform(0.5, 1.6, 102)
            table[0] = npts * 1.7
        st, k, nu = C.c_int(0), C.c_int(0), C.c_double(0)
        harness.h_chi2_bracket(dptr(np.ascontiguousarray(table)), npts, C.byref(st), C.byref(k), C.byref(nu))
        est, ek, enu = _walk_python(table, npts)
        assert (st.value, k.value) == (est, ek)
        if est != 2:
            assert nu.value == enu


def test_lazy_table_gives_the_same_bracket(harness):
    """The table chi2(10^-k) is extended a few decades at a time and only as far as the walk reads
    (csrc/fit.cu k_table_plan / k_table_advance): same bracket as with all 102 entries, fewer systems."""
    rng = np.random.default_rng(11)
    saved = 0
    for trial in range(400):
        npts = int(rng.integers(50, 800))
        kind = trial % 5
        k0 = rng.integers(1, 100)
        lim = int(rng.integers(1, 103)) if trial % 3 else 102
        if kind == 0:
            table = npts * (1.5 - 1.0 / (1 + np.exp(-(np.arange(102) - k0))))
        elif kind == 1:
            table = npts * (1.2 + rng.uniform(0, 1, 102))
        elif kind == 2:
            table = npts * rng.uniform(0.1, 0.5, 102)
        elif kind == 3:
            table = npts * rng.uniform(0.5, 1.6, 102)
            table[0] = npts * 1.7
        else:              # a failed system (NaN) somewhere
            table = npts * (1.5 - 1.0 / (1 + np.exp(-(np.arange(102) - k0))))
            table[int(rng.integers(0, 102))] = np.nan
        table = np.ascontiguousarray(table)
        table[lim:] = table[lim - 1]        # systems with k >= kstar are bit-identical
        for step in (1, 4, 7, 16):
            st, k, nu = C.c_int(0), C.c_int(0), C.c_double(0)
            done = harness.h_chi2_bracket_lazy(dptr(table), npts, step, lim, C.byref(st), C.byref(k), C.byref(nu))
            est, ek, enu = _walk_python(table, npts)
            assert (st.value, k.value) == (est, ek)
            if est != 2:
                assert nu.value == enu
            if est == 0 and enu == npts * 0.6:      # bracket at the first scale factor: nothing past it is touched
                assert done <= min(lim, (ek // step + 1) * step)
            saved += lim - done
    assert saved > 0


def test_nelder_mead_state_machine_replays_scipy(harness):
    """Same abscissae, same minimiser, same nit / nfev / success as scipy's Nelder-Mead (interpolate.py:291)."""
    FN = C.CFUNCTYPE(C.c_double, C.c_double)
    cases = [(lambda a: (a + 24.3) ** 2 + 3.0, -20.0),
             (lambda a: np.cosh(0.3 * (a + 17.77)) * 1e3, -20.0),
             (lambda a: abs(a + 31.4159) ** 1.5, -20.0),
             (lambda a: 5.0 + 0.0 * a, -20.0),                       # flat: converges on the tolerances
             (lambda a: -a * 1e3, -20.0),                            # unbounded: runs into maxiter/maxfev -> failure
             (lambda a: float("nan") if a < -22 else (a + 21.5) ** 2, -20.0),
             (lambda a: (a - 0.3) ** 2, 0.0)]                        # zero start: zdelt branch
    for f, x0 in cases:
        xs_ref = []
        def fr(x):
            xs_ref.append(float(x[0]))
            return f(float(x[0]))
        ref = scipy.optimize.minimize(fr, x0, method="Nelder-Mead")
        xs = np.zeros(512)
        xmin, nfev, nit = C.c_double(0), C.c_int(0), C.c_int(0)
        ok = harness.h_nm_minimize(FN(lambda x: float(f(x))), C.c_double(x0), C.byref(xmin), dptr(xs), C.byref(nfev),
                                   C.byref(nit))
        assert list(xs[:nfev.value]) == xs_ref, (x0, xs[:5], xs_ref[:5])
        assert xmin.value == ref.x[0] and nfev.value == ref.nfev and bool(ok) == bool(ref.success)
        assert nit.value == ref.nit


def test_flattened_ql_is_bit_identical_to_nested(harness):
    """vi_tql_values_flat (the per-lane state machine the GPU runs) == vi_tql_values, bit for bit, on random
    graded tridiagonals, on matrices with exact zeros (splits), and on the real N = 144 systems."""
    from scipy.linalg import lapack
    rng = np.random.default_rng(11)
    mats = []
    for n in (1, 2, 3, 5, 17, 64, 144):
        for _ in range(6):
            d = rng.standard_normal(n) * 10.0 ** rng.uniform(-12, 0, n)
            e = rng.standard_normal(max(n - 1, 0)) * 10.0 ** rng.uniform(-12, 0, max(n - 1, 0))
            if n > 4:
                e[rng.integers(0, n - 1, 2)] = 0.0          # exact splits
            if n > 8:
                d[n // 2:] *= 1e-18                           # graded: many negligible eigenvalues
                e[n // 2:] *= 1e-18
            mats.append((d, e))
    g = load_golden("c1_144")
    ok = np.isfinite(g["value"][0])
    G, y = rp.normal_equations(g["A"][ok], g["error"][0][ok] ** -2, g["value"][0][ok])
    for alpha in (0.0, -10.0, -22.4, -30.0, -60.0):
        X = 0.5 * (G + G.T) + 10.0 ** alpha * g["regs"][0]
        c, d, e, tau, info = lapack.dsytrd(X / np.abs(X).max(), lower=1)
        mats.append((np.ascontiguousarray(d), np.ascontiguousarray(e)))
    for d, e in mats:
        n = d.size
        nrot = C.c_int(0)
        rc = harness.h_tql_flat_identical(n, dptr(np.ascontiguousarray(d)), dptr(np.append(e, 0.0)), C.byref(nrot))
        assert rc == 0, (n, rc)


@pytest.mark.parametrize("n,kind", [(5, "rand"), (27, "rand"), (64, "graded"), (144, "graded"), (144, "rand"), (144, "split"),
                                    (160, "graded")])
def test_wavefront_tape_replay_is_bit_identical(harness, n, kind):
    """vi_wave.h (one warp per system, QL sweeps pipelined across the lanes) against the sequential tape replay of
    vi_tql.h, both directions: the same rotations on the same operands in a different interleaving -> identical bits.
    'split': a tridiagonal with negligible off-diagonals in the middle (short sweeps on sub-blocks, non-nested ranges)."""
    import ctypes as C
    rng = np.random.default_rng(n * 7 + len(kind))
    d = rng.standard_normal(n)
    e = rng.standard_normal(n)
    if kind == "graded":
        d *= 10.0 ** rng.uniform(-12, 0, n)
        e *= 10.0 ** rng.uniform(-12, 0, n)
    if kind == "split":
        e[n // 3] = 1e-300
        e[n // 2] = 0.0
        e[n // 2 + 5] = 1e-40
    g = rng.standard_normal(n)
    ns, nr = C.c_int(0), C.c_int(0)
    fn = harness.h_wave_replay_identical
    fn.restype = C.c_int
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    rc = fn(n, dp(d), dp(e), dp(g), C.byref(ns), C.byref(nr))
    assert rc == 0, (rc, ns.value, nr.value)
    assert 0 < ns.value <= 4 * n + 32 and nr.value >= n - 1
    # per-warp slices of the kernel's shared memory stay 16-byte aligned (the vector is read as doubles)
    assert all(harness.h_wav_bytes(m) % 16 == 0 and harness.h_wav_bytes(m) >= 8 * m + 4 * (4 * m + 33) for m in range(1, 600))
