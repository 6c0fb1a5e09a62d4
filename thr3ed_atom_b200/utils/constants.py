"""Numeric constants and dictionary keys of the render path.

Values and key strings are part of the reference's contract (reference thre3d_atom/utils/constants.py:1-27): the two
epsilons enter the arithmetic of the kernels, the EXTRA_* strings are the keys of ``RenderOut.extra`` and the remaining
strings are keys of saved checkpoints.  They are exported as module attributes under the reference's names.
"""
_DIMENSIONS = dict(NUM_COORD_DIMENSIONS=3, NUM_COLOUR_CHANNELS=3, NUM_RGBA_CHANNELS=4)
_NUMERICS = dict(SEED=42, ZERO_PLUS=1e-10, INFINITY=1e10)  # 1e-10 guards divisions, 1e10 is the last sample's interval
_RENDER_OUT_EXTRA_KEYS = dict(
    EXTRA_DISPARITY="disparity",
    EXTRA_ACCUMULATED_WEIGHTS="accumulated_weight",
    # per-sample debug tensors of the reference's accumulator (never produced by the fused kernels)
    EXTRA_POINT_DENSITIES="point_densities",
    EXTRA_POINT_OCCUPANCIES="point_occupancies",
    EXTRA_SAMPLE_INTERVALS="deltas",
    EXTRA_POINT_WEIGHTS="point_weights",
    EXTRA_POINT_DEPTHS="point_depths",
)
_CHECKPOINT_KEYS = dict(
    CAMERA_BOUNDS="camera_bounds",
    CAMERA_INTRINSICS="camera_intrinsics",
    HEMISPHERICAL_RADIUS="hemispherical_radius",
    EXTRA_INFO="extra_info",
)

for _table in (_DIMENSIONS, _NUMERICS, _RENDER_OUT_EXTRA_KEYS, _CHECKPOINT_KEYS):
    globals().update(_table)
__all__ = [name for _table in (_DIMENSIONS, _NUMERICS, _RENDER_OUT_EXTRA_KEYS, _CHECKPOINT_KEYS) for name in _table]
