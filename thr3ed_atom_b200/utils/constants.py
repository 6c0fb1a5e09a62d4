"""Numeric constants and dictionary keys of the render path.

Values and key strings are part of the reference's contract (reference thre3d_atom/utils/constants.py:1-27): the two
epsilons enter the arithmetic of the kernels, the EXTRA_* strings are the keys of ``RenderOut.extra`` and the remaining
strings are keys of saved checkpoints.  A constant table is what it is: same names, same values as the reference.
"""
from typing import Final

# dimensions
NUM_COORD_DIMENSIONS: Final[int] = 3
NUM_COLOUR_CHANNELS: Final[int] = 3
NUM_RGBA_CHANNELS: Final[int] = 4

# numerics: 1e-10 guards divisions, 1e10 is the last sample's interval (accumulate.py:50-53)
SEED: Final[int] = 42
ZERO_PLUS: Final[float] = 1e-10
INFINITY: Final[float] = 1e10

# keys of RenderOut.extra
EXTRA_DISPARITY: Final[str] = "disparity"
EXTRA_ACCUMULATED_WEIGHTS: Final[str] = "accumulated_weight"
# per-sample debug tensors of the reference's accumulator (never produced by the fused kernels)
EXTRA_POINT_DENSITIES: Final[str] = "point_densities"
EXTRA_POINT_OCCUPANCIES: Final[str] = "point_occupancies"
EXTRA_SAMPLE_INTERVALS: Final[str] = "deltas"
EXTRA_POINT_WEIGHTS: Final[str] = "point_weights"
EXTRA_POINT_DEPTHS: Final[str] = "point_depths"

# keys of the extra-info part of a checkpoint
CAMERA_BOUNDS: Final[str] = "camera_bounds"
CAMERA_INTRINSICS: Final[str] = "camera_intrinsics"
HEMISPHERICAL_RADIUS: Final[str] = "hemispherical_radius"
EXTRA_INFO: Final[str] = "extra_info"
