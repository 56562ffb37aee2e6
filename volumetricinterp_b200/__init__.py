"""volumetricinterp_b200 — B200-native fit/Estimate hot paths of amisr/volumetricinterp.

Same public names as the reference package (volumetricinterp/__init__.py:1-3).
"""
__version__ = "0.1.0"

from .interpolate import Interpolate   # noqa: E402,F401
from .estimate import Estimate         # noqa: E402,F401
