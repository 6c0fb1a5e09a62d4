"""Ray generation and the thin host glue around the render procedure.

Mirrors reference ``thre3d_atom/rendering/volumetric/utils/misc.py``: ``cast_rays`` :12-50 (a CUDA
kernel here), ``flatten_rays`` :53, ``collate_rays`` :60, the ReLU-field density scale :68-78,
synchronous ray/pixel sub-sampling :117-129 and the RenderOut collation helpers :132-163;
``sample_training_ray_batch`` is the fused device-side form of the trainer's batch assembly (trainers.py:281-303).
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


def sample_training_ray_batch(poses: Sequence[CameraPose], camera_intrinsics: CameraIntrinsics, images: Tensor, batch_size: int,
                              tile: Tuple[int, int] = (8, 4), seed: int = None) -> Tuple[Rays, Tensor]:
    """Device-side replacement of the trainer's per-iteration batch assembly (reference modules/trainers.py:281-303):
    ``cast_rays`` for every cached view -> ``collate_rays`` -> ``sample_random_rays_and_pixels_synchronously`` (a ``randperm``
    over all ``V*H*W`` pixels and three gathers).  One kernel draws ``batch_size`` rays and their target pixels straight from
    the poses: nothing of size ``V*H*W`` is generated or permuted.

    ``images``: ``[V, H, W, 3]`` fp32 CUDA tensor (channel-last).  ``tile = (width, height)``: pixels are drawn in whole tiles
    -- ``(8, 4)`` gives every warp of the render kernels one coherent 8x4 pixel tile, ``(1, 1)`` independent pixels (the
    reference's distribution, up to drawing with replacement).  ``seed`` defaults to a draw from torch's CPU generator.
    Returns flat ``Rays [batch, 3]`` and ``pixels [batch, 3]``; rays equal ``cast_rays`` of the same pixel bit for bit."""
    height, width, focal = camera_intrinsics
    device = images.device
    rot = torch.stack([torch.as_tensor(p.rotation, dtype=torch.float32).reshape(3, 3) for p in poses]).to(device)
    trans = torch.stack([torch.as_tensor(p.translation, dtype=torch.float32).reshape(3) for p in poses]).to(device)
    if seed is None:
        seed = int(torch.randint(0, 2**62, (1,), dtype=torch.int64).item())
    origins, directions, pixels, _ = _kernels.sample_ray_batch(rot, trans, images, int(height), int(width), float(focal), int(batch_size),
                                                               tile=tile, seed=seed)
    return Rays(origins, directions), pixels


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
