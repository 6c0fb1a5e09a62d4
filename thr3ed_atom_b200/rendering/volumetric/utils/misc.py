"""Ray generation and the thin host glue around the render procedure.

Mirrors reference ``thre3d_atom/rendering/volumetric/utils/misc.py``: ``cast_rays`` :12-50 (a CUDA
kernel here), ``flatten_rays`` :53, ``collate_rays`` :60, the ReLU-field density scale :68-78,
synchronous ray/pixel sub-sampling :117-129 and the RenderOut collation helpers :132-163.
``ndcize_rays`` is only used by a debug plot in the reference and is out of scope.
"""
from __future__ import annotations

import math
from typing import Sequence, Tuple

import numpy as np
import torch
from torch import Tensor

from thr3ed_atom_b200 import _kernels
from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays, RenderOut
from thr3ed_atom_b200.utils.constants import NUM_COORD_DIMENSIONS
from thr3ed_atom_b200.utils.imaging_utils import CameraIntrinsics, CameraPose


def cast_rays(camera_intrinsics: CameraIntrinsics, pose: CameraPose, device: torch.device = torch.device("cuda")) -> Rays:
    """Pixel-centre pinhole rays ``[H, W, 3]`` for a camera pose, generated on the GPU.

    ``dir = R @ ((x+.5-W/2)/f, -(y+.5-H/2)/f, -1)`` (not normalised), ``origin = t`` broadcast.
    """
    height, width, focal = camera_intrinsics
    origins, directions = _kernels.cast_rays(int(height), int(width), float(focal), pose.rotation, pose.translation, device)
    return Rays(origins.view(height, width, 3), directions.view(height, width, 3))


def flatten_rays(rays: Rays) -> Rays:
    return Rays(rays.origins.reshape(-1, NUM_COORD_DIMENSIONS), rays.directions.reshape(-1, NUM_COORD_DIMENSIONS))


def collate_rays(rays_list: Sequence[Rays]) -> Rays:
    return Rays(torch.cat([r.origins for r in rays_list], dim=0), torch.cat([r.directions for r in rays_list], dim=0))


def compute_expected_density_scale_for_relu_field_grid(grid_world_size: Tuple[float, float, float]) -> float:
    """``(sqrt(27) * 100 / |diagonal|) / 3`` -- 33.33 for the default 3x3x3 world."""
    diagonal = math.sqrt(sum(extent**2 for extent in grid_world_size))
    return ((math.sqrt(3.0**3) * 100.0) / diagonal) / NUM_COORD_DIMENSIONS


def sample_random_rays_and_pixels_synchronously(rays: Rays, pixels: Tensor, sample_size: int) -> Tuple[Rays, Tensor]:
    chosen = torch.randperm(pixels.shape[0], dtype=torch.long, device=pixels.device)[:sample_size]
    return Rays(rays.origins[chosen, :], rays.directions[chosen, :]), pixels[chosen, :]


def collate_rendered_output(rendered_chunks: Sequence[RenderOut]) -> RenderOut:
    keys = list(rendered_chunks[0].extra.keys()) if rendered_chunks else []
    return RenderOut(
        colour=torch.cat([c.colour for c in rendered_chunks], dim=0),
        depth=torch.cat([c.depth for c in rendered_chunks], dim=0),
        extra={k: torch.cat([c.extra[k] for c in rendered_chunks], dim=0) for k in keys},
    )


def reshape_rendered_output(rendered_output: RenderOut, camera_intrinsics: CameraIntrinsics) -> RenderOut:
    shape = (camera_intrinsics.height, camera_intrinsics.width, -1)
    return RenderOut(
        colour=rendered_output.colour.reshape(*shape),
        depth=rendered_output.depth.reshape(*shape),
        extra={k: v.reshape(*shape) for k, v in rendered_output.extra.items()},
    )
