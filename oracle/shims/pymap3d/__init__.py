"""Test-infrastructure shim (NOT product code): stands in for the third-party
`pymap3d` package (requirements.txt:4 of the reference, `>=1.8.0`), which is
absent from this image.  Only `geodetic2ecef` is used on the hot path
(sphharmlag.py:345,351; radbasfun.py:57,253; interpolate.py:422; estimate.py:172).

Restates pymap3d's published WGS84 formula (pymap3d/ecef.py, v2.x/3.x):
    N = a^2 / hypot(a cos(lat), b sin(lat))
    x = (N + h) cos(lat) cos(lon); y = (N + h) cos(lat) sin(lon)
    z = (N (b/a)^2 + h) sin(lat)
parity unpinned: no pymap3d binary/source is available here to pin against.
"""
import numpy as np

_A = 6378137.0
_F = 1.0 / 298.257223563
_B = _A * (1.0 - _F)


def geodetic2ecef(lat, lon, alt, ell=None, deg=True):
    lat = np.asarray(lat, dtype=float)
    lon = np.asarray(lon, dtype=float)
    alt = np.asarray(alt, dtype=float)
    if deg:
        lat = np.radians(lat)
        lon = np.radians(lon)
    n = _A**2 / np.hypot(_A * np.cos(lat), _B * np.sin(lat))
    x = (n + alt) * np.cos(lat) * np.cos(lon)
    y = (n + alt) * np.cos(lat) * np.sin(lon)
    z = (n * (_B / _A) ** 2 + alt) * np.sin(lat)
    return x, y, z
