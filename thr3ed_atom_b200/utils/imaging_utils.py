"""Camera tuples and the small host-side camera math around the renderer.

Mirrors the parts of the reference's ``thre3d_atom/utils/imaging_utils.py`` that the render path and
its callers touch (types :17-30, ``adjust_dynamic_range`` :42-71, ``scale_camera_intrinsics`` :130-138,
``pose_spherical`` :185-191 and the two animation paths :199-234).  The matplotlib-based depth
colouring is presentation code and out of scope.
"""
from __future__ import annotations

import math
from typing import NamedTuple, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor


class CameraIntrinsics(NamedTuple):
    height: int
    width: int
    focal: float


class CameraPose(NamedTuple):
    rotation: Union[np.ndarray, Tensor]  # [3, 3] camera-to-world
    translation: Union[np.ndarray, Tensor]  # [3, 1]


class CameraBounds(NamedTuple):
    near: float
    far: float


def to8b(x: np.ndarray) -> np.ndarray:
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def range_map_coefficients(drange_in: Tuple[float, float], drange_out: Tuple[float, float]) -> Tuple[np.float32, np.float32]:
    """fp32 (scale, bias) of the affine map ``drange_in -> drange_out`` exactly as the reference's
    ``adjust_dynamic_range(slack=True)`` forms them (imaging_utils.py:58-63).  The voxel grid feeds
    these two numbers to the kernels (R3dGrid.norm_scale / norm_bias)."""
    scale = (np.float32(drange_out[1]) - np.float32(drange_out[0])) / (np.float32(drange_in[1]) - np.float32(drange_in[0]))
    bias = np.float32(drange_out[0]) - np.float32(drange_in[0]) * scale
    return np.float32(scale), np.float32(bias)


def adjust_dynamic_range(data, drange_in, drange_out, slack: bool = False):
    """Affine re-mapping of ``data`` between value ranges; ``slack=False`` also clips to the target."""
    if drange_in == drange_out:
        return data
    if slack:
        scale, bias = range_map_coefficients(drange_in, drange_out)
        return data * scale + bias
    lo_i, hi_i = np.float32(drange_in[0]), np.float32(drange_in[1])
    lo_o, hi_o = np.float32(drange_out[0]), np.float32(drange_out[1])
    data = ((data - lo_i) / (hi_i - lo_i) * (hi_o - lo_o)) + lo_o
    return data.clip(drange_out[0], drange_out[1])


def scale_camera_intrinsics(camera_intrinsics: CameraIntrinsics, scale_factor: float = 1.0) -> CameraIntrinsics:
    return CameraIntrinsics(
        height=int(np.ceil(camera_intrinsics.height * scale_factor)),
        width=int(np.ceil(camera_intrinsics.width * scale_factor)),
        focal=camera_intrinsics.focal * scale_factor,
    )


def _homogeneous(rows, device) -> Tensor:
    return torch.tensor(rows, dtype=torch.float32, device=device)


def pose_spherical(yaw: float, pitch: float, radius: float, device=torch.device("cpu")) -> CameraPose:
    """Camera on a sphere looking at the origin: ``Rz(yaw) @ Rx(pitch) @ Tz(radius)``, angles in degrees."""
    y, p = yaw / 180.0 * np.pi, pitch / 180.0 * np.pi
    t_z = _homogeneous([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, radius], [0, 0, 0, 1]], device)
    r_x = _homogeneous([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], device)
    r_z = _homogeneous([[np.cos(y), -np.sin(y), 0, 0], [np.sin(y), np.cos(y), 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], device)
    c2w = r_z @ (r_x @ t_z)
    return CameraPose(rotation=c2w[:3, :3], translation=c2w[:3, 3:])


def get_thre360_animation_poses(hemispherical_radius: float, camera_pitch: float, num_poses: int) -> Sequence[CameraPose]:
    yaws = np.linspace(0, 360, num_poses)[:-1]
    return [pose_spherical(yaw, camera_pitch, hemispherical_radius) for yaw in yaws]


def get_thre360_spiral_animation_poses(
    horizontal_radius_range: Tuple[float, float], vertical_camera_height: float, num_rounds: int, num_poses: int
) -> Sequence[CameraPose]:
    radii = np.linspace(*horizontal_radius_range, num_poses)[:-1]  # last one dropped so the loop closes
    yaws = np.linspace(0, 360 * num_rounds, num_poses)[:-1]
    poses = []
    for yaw, rad in zip(yaws, radii):
        pitch = math.atan(rad / vertical_camera_height) * 180 / math.pi
        poses.append(pose_spherical(yaw, pitch, np.sqrt(rad**2 + vertical_camera_height**2)))
    return poses
