"""Model facade: a 3-D representation + the procedure that renders it + that procedure's configuration.

Keeps the public surface of the reference's ``VolumetricModel`` (thre3d_atom/modules/volumetric_model.py:30-174) and of
``create_volumetric_model_from_saved_model`` (:177-197): callers hold one of these, call ``render_rays`` while training
and ``render`` for whole images, and checkpoint through ``get_save_info``.  It is control flow only -- the work happens in
the render procedure.  Two deliberate differences, both invisible in results:

* ``render`` recognises the fused SH voxel-grid procedure and renders the whole image in ONE launch with in-kernel ray
  generation.  The reference chunks rays (``parallel_rays_chunk_size``) and points (``parallel_points_chunk_size``) only
  because its op-by-op pipeline materialises ``[rays * samples, F + 1]`` tensors; both arguments are accepted and are
  only used for procedures other than the fused one.
* checkpoints are read with ``weights_only=False``: they hold pickled callables and NamedTuples by design
  (reference volumetric_model.py:86-96) and torch >= 2.6 refuses those by default.
"""
from __future__ import annotations

import copy
import dataclasses
from pathlib import Path
from typing import Any, Callable, Dict, Optional, Tuple

import torch
from torch.nn import Module

from thr3ed_atom_b200.rendering.volumetric.render_interface import Rays, RenderOut
from thr3ed_atom_b200.rendering.volumetric.utils import misc as ray_utils
from thr3ed_atom_b200.thre3d_reprs import constants as ckpt
from thr3ed_atom_b200.thre3d_reprs.renderers import (
    RenderConfig,
    RenderProcedure,
    render_sh_voxel_grid,
    render_sh_voxel_grid_camera,
    render_sh_voxel_grid_with_diffuse,
)
from thr3ed_atom_b200.utils.constants import EXTRA_INFO
from thr3ed_atom_b200.utils.imaging_utils import CameraIntrinsics, CameraPose

_DEFAULT_DEVICE = torch.device("cuda" if torch.cuda.is_available() else "cpu")
_CPU = torch.device("cpu")


def _with_overrides(render_config: RenderConfig, overrides: Dict[str, Any]) -> RenderConfig:
    """A deep copy of ``render_config`` with some fields replaced; naming a field it does not have is an error."""
    if not overrides:
        return copy.deepcopy(render_config)
    patched = copy.deepcopy(render_config)
    for name in overrides:
        if not hasattr(patched, name):
            raise ValueError(f"Unknown render configuration field {name} requested for overriding :(")
    for name, value in overrides.items():
        setattr(patched, name, value)
    return patched


class VolumetricModel:
    def __init__(
        self,
        thre3d_repr: Module,
        render_procedure: RenderProcedure,
        render_config: RenderConfig,
        device: torch.device = _DEFAULT_DEVICE,
    ) -> None:
        self._device = device
        self._render_config = render_config
        self._render_procedure = render_procedure
        self._thre3d_repr = thre3d_repr.to(device)

    # ---- read access to the three parts (the representation can be swapped, e.g. after a grid up-scale) ----
    device = property(lambda self: self._device)
    render_config = property(lambda self: self._render_config)
    render_procedure = property(lambda self: self._render_procedure)

    @property
    def thre3d_repr(self) -> Module:
        return self._thre3d_repr

    @thre3d_repr.setter
    def thre3d_repr(self, new_repr: Module) -> None:
        self._thre3d_repr = new_repr

    _update_render_config = staticmethod(_with_overrides)

    # ---- rendering ----
    def render_rays(self, rays: Rays, parallel_points_chunk_size: Optional[int] = None, **kwargs) -> RenderOut:
        """Differentiable render of a flat ray batch; ``kwargs`` override render-config fields for this call only."""
        call_config = _with_overrides(self._render_config, kwargs)
        return self._render_procedure(self._thre3d_repr, rays, call_config, parallel_points_chunk_size)

    def render_rays_with_diffuse(self, rays: Rays, parallel_points_chunk_size: Optional[int] = None, **kwargs) -> Tuple[RenderOut, RenderOut]:
        """``(render_rays(rays), render_rays(rays, render_diffuse=True))`` from one fused march (the pair the reference
        trainer computes with two calls, ``modules/trainers.py:306-330``); both images share the stratified sample
        positions.  Other procedures than the fused SH voxel-grid renderer fall back to the two calls."""
        call_config = _with_overrides(self._render_config, kwargs)
        if self._render_procedure is render_sh_voxel_grid:
            return render_sh_voxel_grid_with_diffuse(self._thre3d_repr, rays, call_config, parallel_points_chunk_size)
        return (self._render_procedure(self._thre3d_repr, rays, call_config, parallel_points_chunk_size),
                self._render_procedure(self._thre3d_repr, rays, _with_overrides(call_config, {"render_diffuse": True}), parallel_points_chunk_size))

    def render(
        self,
        camera_pose: CameraPose,
        camera_intrinsics: CameraIntrinsics,
        parallel_rays_chunk_size: Optional[int] = 32768,
        parallel_points_chunk_size: Optional[int] = None,
        gpu_render: bool = True,
        verbose: bool = False,
        **kwargs,
    ) -> RenderOut:
        """Gradient-free render of the full ``[H, W]`` image seen by a camera; ``kwargs`` override render-config fields.
        ``gpu_render=False`` returns CPU tensors."""
        if self._render_procedure is render_sh_voxel_grid:
            image = render_sh_voxel_grid_camera(self._thre3d_repr, camera_intrinsics, camera_pose, _with_overrides(self._render_config, kwargs))
        else:
            image = self._render_in_ray_chunks(camera_pose, camera_intrinsics, parallel_rays_chunk_size, parallel_points_chunk_size, kwargs)
        if not gpu_render:
            image = image.to(_CPU)
        return ray_utils.reshape_rendered_output(image, camera_intrinsics=camera_intrinsics)

    def _render_in_ray_chunks(self, camera_pose, camera_intrinsics, rays_per_chunk, points_per_chunk, overrides) -> RenderOut:
        """Generic path for user-supplied procedures: cast all rays, render them chunk by chunk without autograd."""
        all_rays = ray_utils.flatten_rays(ray_utils.cast_rays(camera_intrinsics, camera_pose, device=self._device))
        total = len(all_rays)
        step = total if rays_per_chunk is None else rays_per_chunk
        pieces = []
        with torch.no_grad():
            for begin in range(0, total, step):
                pieces.append(self.render_rays(all_rays[begin : begin + step], points_per_chunk, **overrides))
        return ray_utils.collate_rendered_output(pieces)

    # ---- checkpoints ----
    def get_save_info(self, extra_info: Optional[Dict[str, Any]] = None) -> Dict[str, Any]:
        """Everything needed to rebuild this model: representation state + config, the procedure, the config type and values."""
        payload = {
            ckpt.THRE3D_REPR: {
                ckpt.STATE_DICT: self._thre3d_repr.state_dict(),
                ckpt.CONFIG_DICT: self._thre3d_repr.get_save_config_dict(),
            },
            ckpt.RENDER_PROCEDURE: self._render_procedure,
            ckpt.RENDER_CONFIG_TYPE: type(self._render_config),
            ckpt.RENDER_CONFIG: dataclasses.asdict(self._render_config),
        }
        if extra_info is not None:
            payload[EXTRA_INFO] = extra_info
        return payload


def create_volumetric_model_from_saved_model(
    model_path: Path,
    thre3d_repr_creator: Callable[[Dict[str, Any]], Module],
    device: torch.device = _CPU,
) -> Tuple[VolumetricModel, Dict[str, Any]]:
    """Inverse of ``torch.save(vol_mod.get_save_info(extra))``: returns the rebuilt model and the saved extra info."""
    saved = torch.load(model_path, weights_only=False)
    config = saved[ckpt.RENDER_CONFIG_TYPE](**saved[ckpt.RENDER_CONFIG])
    model = VolumetricModel(
        thre3d_repr=thre3d_repr_creator(saved),
        render_procedure=saved[ckpt.RENDER_PROCEDURE],
        render_config=config,
        device=device,
    )
    return model, saved[EXTRA_INFO]
