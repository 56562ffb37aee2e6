"""WGS84 geodetic -> ECEF on the host (numpy).

Restates pymap3d.geodetic2ecef (absent from the image; reference call sites
models/sphharmlag.py:345,351, models/radbasfun.py:57,253, interpolate.py:422,
estimate.py:172).  Host use only: model-centre constants, RBF centres, hull
vertices.  Per-point work on the hot path runs in csrc/vi_math.h.
"""
import numpy as np

WGS84_A = 6378137.0
WGS84_F = 1.0 / 298.257223563
WGS84_B = WGS84_A * (1.0 - WGS84_F)


def geodetic2ecef(lat, lon, alt):
    lat = np.radians(np.asarray(lat, dtype=float))
    lon = np.radians(np.asarray(lon, dtype=float))
    alt = np.asarray(alt, dtype=float)
    n = WGS84_A**2 / np.hypot(WGS84_A * np.cos(lat), WGS84_B * np.sin(lat))
    x = (n + alt) * np.cos(lat) * np.cos(lon)
    y = (n + alt) * np.cos(lat) * np.sin(lon)
    z = (n * (WGS84_B / WGS84_A) ** 2 + alt) * np.sin(lat)
    return x, y, z
