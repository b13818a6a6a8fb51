"""B200-native synthetic tropical-cyclone ensemble integrator (drop-in for the per-year track
generation loop of linjonathan/tropical_cyclone_risk: util/compute.py:64-210)."""
from . import layout, params  # noqa: F401

__all__ = ["layout", "params", "fields", "synth", "engine", "namelist"]
