"""Stand-in for ``matplotlib`` (imported at module level by reference utils/imaging_utils.py:4 and visualizations/static.py:8):
just enough for the hot path's callers -- a "magma" colour map for ``postprocess_depth_map`` (imaging_utils.py:112) and
no-op figure calls for the camera-ray debug plot (static.py:41-79).  Only used when the real package is not installed."""
__version__ = "0.0-standin"
