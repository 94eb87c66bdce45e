"""lightloam_b200 — B200-native Light-LOAM hot path (feature extraction, association + graph vote, LM solve).

The directory name carries a hyphen; import it with importlib:  importlib.import_module("light-loam_b200").
"""
from . import capi, multigpu, synth  # noqa: F401
from .capi import Context, LightLoamError, default_config  # noqa: F401
