"""volumetricinterp_b200 — B200-native fit/Estimate hot paths of amisr/volumetricinterp."""
__version__ = "0.1.0"
