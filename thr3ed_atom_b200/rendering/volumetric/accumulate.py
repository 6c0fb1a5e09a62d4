"""Occupancy model of the accumulator stage.

The fused kernels implement exactly one occupancy model -- the reference's physically based
``alpha = 1 - exp(-sigma * delta)`` (``density2occupancy_pb``, reference
thre3d_atom/rendering/volumetric/accumulate.py:24-28) -- and exactly one tone map
(``torch.sigmoid``).  This function object is the *selector* stored in ``SHVoxGridRenderConfig``
(and pickled into checkpoints by qualified name); ``render_sh_voxel_grid`` checks identity against
it and refuses anything else instead of silently changing semantics.  Calling it evaluates the
formula with torch ops, which is only useful for inspecting values.
"""
import torch
from torch import Tensor


def density2occupancy_pb(densities: Tensor, deltas: Tensor) -> Tensor:
    return 1.0 - torch.exp(-(densities * deltas))
