"""Numeric constants and dictionary keys of the render path.

Values and key strings are part of the reference's contract (thre3d_atom/utils/constants.py:1-27):
ZERO_PLUS / INFINITY enter the arithmetic of the kernels, the EXTRA_* strings are the keys of
``RenderOut.extra`` and the CAMERA_* / EXTRA_INFO strings are keys of saved checkpoints.
"""
NUM_COORD_DIMENSIONS = 3
NUM_COLOUR_CHANNELS = 3
NUM_RGBA_CHANNELS = 4

SEED = 42
ZERO_PLUS = 1e-10
INFINITY = 1e10

EXTRA_DISPARITY = "disparity"
EXTRA_ACCUMULATED_WEIGHTS = "accumulated_weight"
EXTRA_POINT_DENSITIES = "point_densities"
EXTRA_POINT_OCCUPANCIES = "point_occupancies"
EXTRA_SAMPLE_INTERVALS = "deltas"
EXTRA_POINT_WEIGHTS = "point_weights"
EXTRA_POINT_DEPTHS = "point_depths"

CAMERA_BOUNDS = "camera_bounds"
CAMERA_INTRINSICS = "camera_intrinsics"
HEMISPHERICAL_RADIUS = "hemispherical_radius"

EXTRA_INFO = "extra_info"
