"""Synthetic AMISR-shaped inputs.

The reference ships no sample data file (its FILENAME, example_config.ini:9, is
not in the repository), so every benchmark/test input is generated here with
fixed seeds.  Shapes follow the AMISR fitted-file layout that
`Interpolate.read_datafile` consumes (reference interpolate.py:582-667):
geometry is (nbeams, ngates), data is (nrecords, nbeams, ngates).

Pure numpy; no GPU, no oracle import.  The data values are produced from a
caller-supplied design matrix `A` (so that the chi^2 = nu root the reference
searches for exists); the caller decides who evaluates the basis (the CUDA
kernel in bench.py, the oracle in the golden-vector script).
"""
import numpy as np

WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
WGS84_B = WGS84_A * (1.0 - WGS84_F)

RADAR_LAT = 74.73     # RISR-N like site
RADAR_LON = 265.09 - 360.0
LATCP = 78.0
LONCP = 262.0


def _geodetic2ecef(lat, lon, alt):
    lat = np.radians(lat)
    lon = np.radians(lon)
    n = WGS84_A**2 / np.hypot(WGS84_A * np.cos(lat), WGS84_B * np.sin(lat))
    return ((n + alt) * np.cos(lat) * np.cos(lon),
            (n + alt) * np.cos(lat) * np.sin(lon),
            (n * (WGS84_B / WGS84_A) ** 2 + alt) * np.sin(lat))


def _ecef2geodetic(x, y, z):
    """Bowring iteration; accuracy far below what a synthetic geometry needs."""
    e2 = 1.0 - (WGS84_B / WGS84_A) ** 2
    lon = np.degrees(np.arctan2(y, x))
    p = np.hypot(x, y)
    lat = np.arctan2(z, p * (1.0 - e2))
    for _ in range(8):
        n = WGS84_A / np.sqrt(1.0 - e2 * np.sin(lat) ** 2)
        alt = p / np.cos(lat) - n
        lat = np.arctan2(z, p * (1.0 - e2 * n / (n + alt)))
    n = WGS84_A / np.sqrt(1.0 - e2 * np.sin(lat) ** 2)
    alt = p / np.cos(lat) - n
    return np.degrees(lat), lon, alt


def make_geometry(nbeams, ngates, seed=0, nan_alt_frac=0.05):
    """Beam/gate geometry: beam 0 vertical, others az~U[0,360), el~U[35,90];
    gates linearly spaced 100-700 km in slant range.  Returns (lat, lon, alt)
    each (nbeams, ngates) in deg/deg/m with ~nan_alt_frac gates NaN-padded in
    all three (AMISR files pad unused gates with NaN)."""
    rng = np.random.default_rng(seed)
    az = np.radians(rng.uniform(0.0, 360.0, nbeams))
    el = np.radians(rng.uniform(35.0, 90.0, nbeams))
    az[0], el[0] = 0.0, np.pi / 2
    rng_m = np.linspace(100e3, 700e3, ngates)
    e = np.cos(el)[:, None] * np.sin(az)[:, None] * rng_m[None, :]
    n = np.cos(el)[:, None] * np.cos(az)[:, None] * rng_m[None, :]
    u = np.sin(el)[:, None] * rng_m[None, :]
    lat0, lon0 = np.radians(RADAR_LAT), np.radians(RADAR_LON)
    x0, y0, z0 = _geodetic2ecef(RADAR_LAT, RADAR_LON, 0.0)
    x = x0 - np.sin(lon0) * e - np.sin(lat0) * np.cos(lon0) * n + np.cos(lat0) * np.cos(lon0) * u
    y = y0 + np.cos(lon0) * e - np.sin(lat0) * np.sin(lon0) * n + np.cos(lat0) * np.sin(lon0) * u
    z = z0 + np.cos(lat0) * n + np.sin(lat0) * u
    lat, lon, alt = _ecef2geodetic(x, y, z)
    lon = np.where(lon < 0, lon + 360.0, lon)
    if nan_alt_frac > 0:
        bad = rng.uniform(size=lat.shape) < nan_alt_frac
        bad[0, :] = False
        lat = np.where(bad, np.nan, lat)
        lon = np.where(bad, np.nan, lon)
        alt = np.where(bad, np.nan, alt)
    return lat, lon, alt


def true_coeffs(nbasis, maxl=None, seed=1, n_terms=5, scale=1e11):
    """A few low-order nonzero coefficients at ~`scale`."""
    rng = np.random.default_rng(seed)
    c = np.zeros(nbasis)
    idx = np.arange(min(n_terms, nbasis))
    if maxl is not None and nbasis > maxl * maxl:
        # first term of the first radial orders plus two horizontal terms
        idx = np.array([0, maxl * maxl, 1, 2, 3][:n_terms]) % nbasis
    c[idx] = scale * rng.uniform(0.5, 1.5, idx.size) * rng.choice([-1.0, 1.0], idx.size)
    c[idx[0]] = abs(c[idx[0]]) * 3.0
    return c


def structured_coeffs(nbasis, maxl, n_terms=10, amp=0.3, seed=1, scale=1e11):
    """Like true_coeffs, with horizontal structure beyond the first five terms (degrees l <= 2 of the first radial
    order, a few terms of the next two) at `amp` of their amplitude.  With the curvature regulariser this puts
    chi2(lambda = 1) well above the gate count while the unregularised chi2 stays near noise_scale^2 of it, so the
    reference's chi2 = nu search (interpolate.py:173-218) finds a root for (almost) every record instead of ending
    "too smooth" (lambda = 0) or without a bracket (NaN record)."""
    rng = np.random.default_rng(seed)
    L2 = maxl * maxl
    order = [0, L2, 1, 2, 3, 4, 5, 6, 7, 8, L2 + 1, L2 + 2, L2 + 3, 2 * L2, 2 * L2 + 1]
    idx = np.array(order[:n_terms]) % nbasis
    c = np.zeros(nbasis)
    c[idx] = scale * rng.uniform(0.5, 1.5, idx.size) * rng.choice([-1.0, 1.0], idx.size)
    c[idx[5:]] *= amp
    c[idx[0]] = abs(c[idx[0]]) * 3.0
    return c


def make_records(A, nrecords, seed=2, c_true=None, bad_frac=0.10, drift=0.05,
                 maxl=None, noise_scale=1.0):
    """Records generated from the model itself: d = A c + sigma N(0,1) with
    sigma = clip(0.05|d| + 2e10, 1.1e10, 9e12) (inside the default ERRLIM); the
    noise actually added is noise_scale*sigma.

    A: (P, N) design matrix at the P valid (non-NaN-altitude) points.
    Returns value, error (nrecords, P) with ~bad_frac of the gates set to NaN
    per record (what the reference's quality filter produces,
    interpolate.py:652-657), and the per-record true coefficients."""
    rng = np.random.default_rng(seed)
    P, N = A.shape
    if c_true is None:
        c_true = true_coeffs(N, maxl=maxl)
    value = np.empty((nrecords, P))
    error = np.empty((nrecords, P))
    ctrue = np.empty((nrecords, N))
    for r in range(nrecords):
        c = c_true * (1.0 + drift * rng.standard_normal(N))
        d0 = A @ c
        sigma = np.clip(0.05 * np.abs(d0) + 2e10, 1.1e10, 9e12)
        d = d0 + noise_scale * sigma * rng.standard_normal(P)
        bad = rng.uniform(size=P) < bad_frac
        d[bad] = np.nan
        sigma[bad] = np.nan
        value[r], error[r], ctrue[r] = d, sigma, c
    return value, error, ctrue


def make_unixtime(nrecords, t0=1480286700.0, dt=60.0):
    """(nrecords, 2) start/end seconds, 1-minute records starting 2016-11-27T22:45."""
    start = t0 + dt * np.arange(nrecords)
    return np.stack([start, start + dt], axis=1)


def make_fitinfo(shape, seed=3):
    """FitInfo arrays that pass the default filter: chi2~U(0.2,5), fitcode in 1..4."""
    rng = np.random.default_rng(seed)
    return rng.uniform(0.2, 5.0, shape), rng.integers(1, 5, shape).astype(np.int64)


def flatten_valid(lat, lon, alt):
    """Drop NaN-altitude gates exactly like interpolate.py:660-664: returns the
    flattened (P,) coordinates and the boolean keep-mask over nbeams*ngates."""
    la, lo, al = lat.ravel(), lon.ravel(), alt.ravel()
    keep = np.isfinite(al)
    return la[keep], lo[keep], al[keep], keep


def write_amisr_file(filename, nbeams, ngates, nrecords, A_of=None, seed=0, noise_scale=1.0):
    """Write a synthetic AMISR fitted file with the layout `Interpolate.read_datafile` expects
    (reference interpolate.py:605-632).  `A_of(lat, lon, alt) -> (P, N)` is the
    design-matrix evaluator used to generate the densities (GPU kernel or oracle, caller's choice).
    Returns the arrays written (for tests)."""
    from . import h5lite
    lat2, lon2, alt2 = make_geometry(nbeams, ngates, seed=seed)
    lat, lon, alt, keep = flatten_valid(lat2, lon2, alt2)
    A = A_of(lat, lon, alt)
    value, error, _ = make_records(A, nrecords, seed=seed + 1, noise_scale=noise_scale, bad_frac=0.0)
    ne = np.full((nrecords, nbeams * ngates), np.nan)
    dne = np.full((nrecords, nbeams * ngates), np.nan)
    ne[:, keep], dne[:, keep] = value, error
    chi2, fitcode = make_fitinfo((nrecords, nbeams, ngates), seed=seed + 2)
    rng = np.random.default_rng(seed + 3)
    # ~10 % of the gates fail the quality filter: bad fit code, chi2 or error out of range
    bad = rng.uniform(size=chi2.shape)
    fitcode[bad < 0.04] = 5
    chi2[(bad >= 0.04) & (bad < 0.07)] = 25.0
    dne3 = dne.reshape(nrecords, nbeams, ngates)
    dne3[(bad >= 0.07) & (bad < 0.10)] = 5e13
    utime = make_unixtime(nrecords)
    arrays = {"/Time/UnixTime": utime, "/Geomag/Altitude": alt2, "/Geomag/Latitude": lat2, "/Geomag/Longitude": lon2,
              "/FittedParams/Ne": ne.reshape(nrecords, nbeams, ngates), "/FittedParams/dNe": dne3,
              "/FittedParams/FitInfo/chi2": chi2, "/FittedParams/FitInfo/fitcode": fitcode,
              "/FittedParams/IonMass": np.array([16.0, 32.0, 30.0])}
    with h5lite.Writer(filename, pytables_attrs=False) as h5:
        for k, v in arrays.items():
            h5.array(k, v)
    return arrays
