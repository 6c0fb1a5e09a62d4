"""Boundary data types of the volumetric render path.

Same names, fields and shape contracts as the reference's
``thre3d_atom/rendering/volumetric/render_interface.py`` (``Rays`` :13-44, ``RenderOut`` :47-83,
``SampledPointsOnRays`` :86-94, ``render`` :103-134), because callers construct and destructure
these objects directly.  ``render`` is kept as the generic three-stage driver for API
compatibility; the SH voxel-grid procedure does not go through it -- it runs the three stages as
one fused CUDA kernel (see ``thre3d_reprs/renderers.py``).
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Dict, NamedTuple, Optional

import torch
from torch import Tensor

from thr3ed_atom_b200.utils.constants import NUM_COLOUR_CHANNELS, NUM_COORD_DIMENSIONS
from thr3ed_atom_b200.utils.imaging_utils import CameraBounds

ExtraInfo = Dict[str, Any]


@dataclasses.dataclass
class Rays:
    origins: Tensor  # [..., 3]
    directions: Tensor  # [..., 3]; not necessarily unit length

    def __post_init__(self):
        assert self.origins.shape == self.directions.shape, "ray-origins and ray-directions are incompatible :("
        assert self.origins.shape[-1] == NUM_COORD_DIMENSIONS, "only 3D rays are supported"

    def __getitem__(self, item) -> "Rays":
        return Rays(self.origins[item, :], self.directions[item, :])

    def __len__(self) -> int:
        return len(self.origins)

    def to(self, device: torch.device) -> "Rays":
        return Rays(self.origins.to(device), self.directions.to(device))


@dataclasses.dataclass
class RenderOut:
    colour: Tensor  # [..., 3]
    depth: Tensor  # [..., 1]
    extra: Optional[ExtraInfo] = None

    def __post_init__(self):
        assert self.colour.shape[:-1] == self.depth.shape[:-1], "rendered colour maps and depth maps are shape-incompatible"
        assert self.colour.shape[-1] == NUM_COLOUR_CHANNELS, "only RGB colours are possible"
        assert self.depth.shape[-1] == 1, "depth map should only have 1 dimensional data channel"
        if self.extra is None:
            self.extra = {}

    def _map(self, fn) -> "RenderOut":
        return RenderOut(fn(self.colour), fn(self.depth), {k: fn(v) for k, v in self.extra.items()})

    def detach(self) -> "RenderOut":
        return self._map(lambda t: t.detach())

    def to(self, device: torch.device) -> "RenderOut":
        return self._map(lambda t: t.to(device))


class SampledPointsOnRays(NamedTuple):
    points: Tensor  # [N, S, 3]
    depths: Tensor  # [N, S]


ProcessedPointsOnRays = SampledPointsOnRays

RaySamplerFunction = Callable[[Rays, CameraBounds, int], SampledPointsOnRays]
PointProcessorFunction = Callable[[SampledPointsOnRays, Rays], ProcessedPointsOnRays]
AccumulatorFunction = Callable[[ProcessedPointsOnRays, Rays], RenderOut]


def assert_flat_rays(rays: Rays) -> None:
    assert (
        len(rays.origins.shape) == len(rays.directions.shape) == 2
    ), "Please note that the RENDER interface only works with FLAT RAYS!"


def render(
    rays: Rays,
    camera_bounds: CameraBounds,
    num_samples: int,
    sampler_fn: RaySamplerFunction,
    point_processor_fn: PointProcessorFunction,
    accumulator_fn: AccumulatorFunction,
) -> RenderOut:
    """sampler -> point processor -> accumulator on flat rays (user-supplied stages)."""
    assert_flat_rays(rays)
    return accumulator_fn(point_processor_fn(sampler_fn(rays, camera_bounds, num_samples), rays), rays)
